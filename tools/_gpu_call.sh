set -x
for v in base m6 m8; do
  if [ $v = base ]; then unset B2F_LIB; else export B2F_LIB=$PWD/tools/_build/$v/libb200fem.so; fi
  echo "=== $v" >> gpurun_out/r02_tpe_minb.log
  python tools/opbench.py --bp 1 --p 1 --n 270 --variants 0 --reps 20 >> gpurun_out/r02_tpe_minb.log 2>&1
done
grep -v "^BP" gpurun_out/r02_tpe_minb.log
