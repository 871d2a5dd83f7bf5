set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -3 gpurun_out/r02_smoke.log
date +%s > gpurun_out/t0
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_driver_ref.json 2> gpurun_out/r02_driver_ref.err; date +%s > gpurun_out/t1
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_driver_n1.json 2> gpurun_out/r02_driver_n1.err; date +%s > gpurun_out/t2
echo "ref arm s: $(( $(cat gpurun_out/t1) - $(cat gpurun_out/t0) ))  b200 arm s: $(( $(cat gpurun_out/t2) - $(cat gpurun_out/t1) ))"
cut -c1-700 gpurun_out/r02_driver_ref.json; tail -3 gpurun_out/r02_driver_ref.err
cut -c1-300 gpurun_out/r02_driver_n1.json; tail -3 gpurun_out/r02_driver_n1.err
