set -x
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_final3_n1.json 2> gpurun_out/r02_final3_n1.err
python -c "
import json; l=json.loads(open('gpurun_out/r02_final3_n1.json').read().strip().splitlines()[-1]); print('N=1', round(l['value'],2), round(l['e2e']['value'],2), round(l['roofline']['frac'],3), l['clocks']['sm_mhz'])"
tail -3 gpurun_out/r02_final3_n1.err
