set -x
nvidia-smi -L
python -m pytest tests/test_gpu_parallel.py -m gpu -x -q > gpurun_out/r02_gputest_parallel_2gpu.log 2>&1; tail -8 gpurun_out/r02_gputest_parallel_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err
tail -c 2500 gpurun_out/r02_bench_2gpu.json; tail -5 gpurun_out/r02_bench_2gpu.err
B2F_HALO=nccl python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r02_bench_2gpu_nccl.json 2> gpurun_out/r02_bench_2gpu_nccl.err
tail -c 1200 gpurun_out/r02_bench_2gpu_nccl.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --bp 5 --steps 2 --warmup 3 > gpurun_out/r02_bench_2gpu_bp5.json 2> gpurun_out/r02_bench_2gpu_bp5.err
tail -c 1500 gpurun_out/r02_bench_2gpu_bp5.json; tail -3 gpurun_out/r02_bench_2gpu_bp5.err
