set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_vector.py -m gpu -x -q -k "p7 or p8 or checksum or large_properties" > gpurun_out/r02_gputest_alias.log 2>&1; tail -4 gpurun_out/r02_gputest_alias.log
for c in "3 7 39" "3 8 34" "5 8 34"; do
  set -- $c
  python tools/opbench.py --bp $1 --p $2 --n $3 --variants c0,c1 --reps 20 >> gpurun_out/r02_alias.log 2>&1
done
cat gpurun_out/r02_alias.log
