set -x
for c in "5 7 66" "1 4 92" "3 4 116" "3 4 133"; do
  set -- $c
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:pa_apply_3d_eo -s 3 -c 1 --csv --log-file gpurun_out/r02_traffic_bp$1_p$2_n$3.csv python tools/opbench.py --bp $1 --p $2 --n $3 --reps 2 > gpurun_out/r02_traffic_bp$1_p$2_n$3.log 2>&1
  tail -4 gpurun_out/r02_traffic_bp$1_p$2_n$3.csv
done
