#!/bin/bash
# GPU call: programmatic dependent launch of gn_step (DGPMP2_PDL=2) -- correctness under PDL, then A/B timing.
set -u
O=gpurun_out/call4; mkdir -p $O
t0=$(date +%s)
DGPMP2_PDL=2 timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_backward.py -m gpu -q -x > $O/tests_pdl.txt 2>&1; echo "pdl tests rc=$? $(( $(date +%s)-t0 ))s" | tee -a $O/log.txt
tail -5 $O/tests_pdl.txt
run() { echo -n "$1: " | tee -a $O/ab.txt; env $2 timeout 60 python scratch/graph_time.py $3 $4 2>&1 | tail -1 | tee -a $O/ab.txt; }
for i in 1 2; do
  run plain DGPMP2_PDL=1 1024 64
  run pdl   DGPMP2_PDL=2 1024 64
done
run plain DGPMP2_PDL=1 1 64
run pdl   DGPMP2_PDL=2 1 64
run plain DGPMP2_PDL=1 1024 128
run pdl   DGPMP2_PDL=2 1024 128
run plain DGPMP2_PDL=1 8192 64
run pdl   DGPMP2_PDL=2 8192 64
echo "done $(( $(date +%s)-t0 ))s" | tee -a $O/log.txt
