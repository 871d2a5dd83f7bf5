set -x
date +%s > gpurun_out/t0
python bench.py > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err
date +%s > gpurun_out/t1
tail -c 3000 gpurun_out/r02_bench_full.json; tail -5 gpurun_out/r02_bench_full.err
echo "bench seconds: $(( $(cat gpurun_out/t1) - $(cat gpurun_out/t0) ))"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_refarm.json 2>&1; cat gpurun_out/r02_bench_refarm.json | cut -c1-600
