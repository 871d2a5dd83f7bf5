set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 > gpurun_out/r02_bench_8gpu_b.json 2> gpurun_out/r02_bench_8gpu_b.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 4 > gpurun_out/r02_bench_4gpu_b.json 2> gpurun_out/r02_bench_4gpu_b.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus 8 --bp 5 > gpurun_out/r02_bench_8gpu_bp5_b.json 2> gpurun_out/r02_bench_8gpu_bp5_b.err
for f in 8gpu_b 4gpu_b 8gpu_bp5_b; do python -c "
import json; l=json.loads(open('gpurun_out/r02_bench_$f.json').read().strip().splitlines()[-1]); print('$f', round(l['value'],2), round(l['e2e']['value'],2), round(l['roofline']['frac'],3), l['parity_check']['ok'], l['config']['setup_seconds'])"; done
tail -2 gpurun_out/r02_bench_8gpu_b.err
