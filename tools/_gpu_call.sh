set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "p5 or p6 or p7 or checksum" > gpurun_out/r02_gputest_alias6.log 2>&1; tail -3 gpurun_out/r02_gputest_alias6.log
for c in "3 5 54" "5 5 54" "5 6 45" "3 7 39"; do
  set -- $c
  python tools/opbench.py --bp $1 --p $2 --n $3 --variants c0 --reps 20 >> gpurun_out/r02_alias6.log 2>&1
done
cat gpurun_out/r02_alias6.log
