#!/bin/bash
# bench.py conditions (rotating input sets > L2): plain launches vs programmatic dependent launch, alternating
set -u
O=gpurun_out/call5; mkdir -p $O
for i in 1 2 3; do
  for m in 1 2; do
    DGPMP2_PDL=$m timeout 60 python bench.py --steps 10000 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('PDL=$m  us/step %.3f  value %.4g  e2e %.4g  clocks %s' % (d['ms_per_step']*1e3, d['value'], d['e2e']['value'], d['clocks']))" | tee -a $O/bench_ab.txt
  done
done
