set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_5.log 2>&1; tail -4 gpurun_out/r02_gputest_5.log
python bench.py --bp 5 --steps 5 --warmup 3 --no-sweep > gpurun_out/r02_bench_bp5_c5.json 2> gpurun_out/r02_bench_bp5_c5.err; cut -c1-400 gpurun_out/r02_bench_bp5_c5.json; tail -3 gpurun_out/r02_bench_bp5_c5.err
python bench.py --bp 1 --steps 5 --warmup 3 --no-sweep > gpurun_out/r02_bench_bp1_c2.json 2> gpurun_out/r02_bench_bp1_c2.err; cut -c1-400 gpurun_out/r02_bench_bp1_c2.json; tail -3 gpurun_out/r02_bench_bp1_c2.err
