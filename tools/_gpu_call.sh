set -x
for v in base v1 v4 v5; do
  if [ $v = base ]; then unset B2F_LIB; else export B2F_LIB=$PWD/tools/_build/$v/libb200fem.so; fi
  echo "=== $v" >> gpurun_out/r02_mass_variants.log
  VAR=0; [ $v = v4 ] && VAR=0,persist; [ $v = v5 ] && VAR=c0
  python tools/opbench.py --sweep 1 --variants $VAR --reps 20 >> gpurun_out/r02_mass_variants.log 2>&1
done
grep -v "^BP" gpurun_out/r02_mass_variants.log
