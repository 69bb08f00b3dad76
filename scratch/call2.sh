#!/bin/bash
# GPU call 2: row-per-lane tail (DGPMP2_TAILMODE=2): correctness, then A/B timing against the block-Thomas tail.
set -u
O=gpurun_out/call2; mkdir -p $O
t0=$(date +%s)
DGPMP2_TAILMODE=2 timeout 200 python -m pytest tests/test_gpu_schedules.py tests/test_gpu_parity.py -m gpu -q -x > $O/tests_rows.txt 2>&1; echo "rows tests rc=$? $(( $(date +%s)-t0 ))s" | tee -a $O/log.txt
tail -12 $O/tests_rows.txt
run() { echo -n "$1: " | tee -a $O/ab.txt; env $2 timeout 60 python scratch/graph_time.py $3 $4 2>&1 | tail -1 | tee -a $O/ab.txt; }
for i in 1 2; do
  run block DGPMP2_TAILMODE=1 1024 64
  run rows  DGPMP2_TAILMODE=2 1024 64
done
run block DGPMP2_TAILMODE=1 1 64
run rows  DGPMP2_TAILMODE=2 1 64
run block DGPMP2_TAILMODE=1 1024 128
run rows  DGPMP2_TAILMODE=2 1024 128
run rows-tail8 "DGPMP2_TAILMODE=2 DGPMP2_TAIL=8" 1024 64
run rows-tail2 "DGPMP2_TAILMODE=2 DGPMP2_TAIL=2" 1024 64
echo "done $(( $(date +%s)-t0 ))s" | tee -a $O/log.txt
