#!/bin/bash
mkdir -p gpurun_out/r3p; cd /root/repo; O=gpurun_out/r3p
python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > $O/pytest.log
python bench.py > $O/bench.json 2> $O/bench.err
timeout 600 ncu --set full --clock-control none -k regex:hinge_kernel -s 3 -c 1 -o $O/hinge python scratch/r2_k1.py 32768 128 5 > $O/k1.log 2>&1
python scratch/ncu_summary.py $O/hinge.ncu-rep $O/hinge_full_summary.csv hinge_kernel > /dev/null 2>&1
rm -f $O/hinge.ncu-rep
cat $O/pytest.log; python - <<'P'
import json
d=json.loads(open('gpurun_out/r3p/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline_k1']['kernel_us'], d['roofline_k1']['frac'], d['clocks'])
P
grep "dram__bytes\|time_duration\|warps_active\|registers" $O/hinge_full_summary.csv
