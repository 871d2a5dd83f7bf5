set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_4.log 2>&1; tail -4 gpurun_out/r02_gputest_4.log
# launch list of the bench command (short run, no baselines / table)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-sweep --no-anchor --no-parity > gpurun_out/r02_ncu_bench.log 2>&1
tail -2 gpurun_out/r02_ncu_bench.log | cut -c1-300
# full captures: swizzled Q=8 kernels, the headline kernel, the diagonal kernel
for c in "3 6 45" "5 7 39" "3 4 96"; do
  set -- $c
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:pa_apply_3d_eo -s 3 -c 1 -f -o gpurun_out/r02_prof2_bp$1_p$2 python tools/opbench.py --bp $1 --p $2 --n $3 --reps 2 > gpurun_out/r02_prof2_bp$1_p$2.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pa_diag_3d -s 2 -c 1 -f -o gpurun_out/r02_prof2_diag python tools/diagbench.py --n 96 --reps 1 > gpurun_out/r02_prof2_diag.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
