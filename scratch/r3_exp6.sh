#!/bin/bash
mkdir -p gpurun_out/r3l; cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r3l/pytest.log
python scratch/r3_cold.py 64 1024 > gpurun_out/r3l/cold.json 2> gpurun_out/r3l/err.txt
python scratch/r3_zerocopy.py > gpurun_out/r3l/zerocopy.json 2>> gpurun_out/r3l/err.txt
python bench.py --steps 2000 --warmup 20 --no-cpu-baseline > gpurun_out/r3l/bench.json 2>> gpurun_out/r3l/err.txt
cat gpurun_out/r3l/pytest.log gpurun_out/r3l/cold.json gpurun_out/r3l/zerocopy.json; python - <<'P'
import json
d=json.loads(open('gpurun_out/r3l/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], {k:v['value'] for k,v in d['config'].items() if k.startswith('e2e')})
print({k:(v.get('us_per_step')) for k,v in d['config']['extras'].items() if isinstance(v,dict)}, d['roofline_k1']['kernel_us'])
P
tail -n 3 gpurun_out/r3l/err.txt
