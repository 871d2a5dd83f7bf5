set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_6.log 2>&1; tail -3 gpurun_out/r02_gputest_6.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_2gpu_b.json 2> gpurun_out/r02_bench_2gpu_b.err
python -c "
import json; l=json.loads(open('gpurun_out/r02_bench_2gpu_b.json').read().strip().splitlines()[-1]); print('N=2', round(l['value'],2), round(l['e2e']['value'],2), l['parity_check']['ok'])"
tail -3 gpurun_out/r02_bench_2gpu_b.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-sweep --no-anchor > gpurun_out/r02_bench_1gpu_b.json 2>/dev/null
python -c "
import json; l=json.loads(open('gpurun_out/r02_bench_1gpu_b.json').read().strip().splitlines()[-1]); print('N=1', round(l['value'],2), round(l['e2e']['value'],2), round(l['roofline']['frac'],3), l['gpu_launches'])"
