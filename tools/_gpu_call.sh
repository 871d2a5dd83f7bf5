set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "thread_per_element or checksum" > gpurun_out/r02_gputest_tpe3.log 2>&1; tail -4 gpurun_out/r02_gputest_tpe3.log
python tools/opbench.py --bp 1 --p 2 --n 135 --variants notpe,tpe --reps 20 > gpurun_out/r02_tpe3.log 2>&1
cat gpurun_out/r02_tpe3.log
