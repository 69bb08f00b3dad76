#!/bin/bash
mkdir -p gpurun_out/r3e; cd /root/repo
python -m pytest tests/test_gpu_host_step.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r3e/pytest.log
python bench.py --steps 2000 --warmup 20 > gpurun_out/r3e/bench.json 2> gpurun_out/r3e/bench.err
DGPMP2_HOST_INPLACE_IO=2 python bench.py --steps 200 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r3e/bench_io_inplace.json 2>> gpurun_out/r3e/bench.err
ncu --query-metrics 2>/dev/null | grep -i pcie > gpurun_out/r3e/pcie_metrics.txt
cat gpurun_out/r3e/pytest.log; python - <<'P'
import json
for f in ('bench.json','bench_io_inplace.json'):
    try:
        d=json.loads(open('gpurun_out/r3e/'+f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['e2e'], {k:v['value'] for k,v in d['config'].items() if k.startswith('e2e')}, d['config']['host'])
        print({k:(v.get('us_per_step') if isinstance(v,dict) else v) for k,v in d['config']['extras'].items()})
        print(d['roofline']['frac'], d.get('roofline_k1'))
    except Exception as e: print(f, 'ERR', e)
P
tail -n 5 gpurun_out/r3e/bench.err; head -20 gpurun_out/r3e/pcie_metrics.txt
