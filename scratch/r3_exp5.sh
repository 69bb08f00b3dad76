#!/bin/bash
mkdir -p gpurun_out/r3f; cd /root/repo
python -m pytest tests/test_gpu_host_step.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r3f/pytest.log
python bench.py --steps 200 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r3f/bench.json 2> gpurun_out/r3f/bench.err
timeout 300 ncu --metrics pcie__read_bytes.sum,pcie__write_bytes.sum,gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:gn_step_kernel -c 6 --csv --log-file gpurun_out/r3f/pcie.csv python scratch/r3_zerocopy.py > gpurun_out/r3f/zc_under_ncu.log 2>&1
cat gpurun_out/r3f/pytest.log; python - <<'P'
import json
d=json.loads(open('gpurun_out/r3f/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], {k:v['value'] for k,v in d['config'].items() if k.startswith('e2e')})
P
tail -n 5 gpurun_out/r3f/bench.err; grep -v "^==" gpurun_out/r3f/pcie.csv | cut -d, -f5,13- | cut -c1-30,200- | head -30
