set -x
B2F_TPE=1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mass3d_p1 or mass3d_p2 or checksum or large_properties or end_to_end" > gpurun_out/r02_gputest_tpe.log 2>&1; tail -4 gpurun_out/r02_gputest_tpe.log
python tools/opbench.py --bp 1 --p 1 --n 270 --variants notpe,tpe --reps 20 > gpurun_out/r02_tpe.log 2>&1
python tools/opbench.py --bp 1 --p 2 --n 135 --variants notpe,tpe --reps 20 >> gpurun_out/r02_tpe.log 2>&1
cat gpurun_out/r02_tpe.log
