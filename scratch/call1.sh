#!/bin/bash
# GPU call 1: head tests, A/B timing of the step kernel (pre-head library vs current), then the full gpu suite.
set -u
O=gpurun_out/call1; mkdir -p $O
t0=$(date +%s)
timeout 240 python -m pytest tests/test_gpu_head.py -m gpu -q -x > $O/head_tests.txt 2>&1; echo "head tests rc=$? $(( $(date +%s)-t0 ))s" | tee -a $O/log.txt
tail -15 $O/head_tests.txt
for i in 1 2 3; do
  DGPMP2_LIB=$PWD/scratch/lib_base.so timeout 60 python scratch/graph_time.py 1024 64 2>&1 | tail -1 | sed 's/^/base: /' | tee -a $O/ab.txt
  timeout 60 python scratch/graph_time.py 1024 64 2>&1 | tail -1 | sed 's/^/head: /' | tee -a $O/ab.txt
done
echo "ab done $(( $(date +%s)-t0 ))s" | tee -a $O/log.txt
timeout 400 python -m pytest tests -m gpu -q -x --durations=8 > $O/gpu_tests.txt 2>&1; echo "gpu tests rc=$? $(( $(date +%s)-t0 ))s" | tee -a $O/log.txt
tail -14 $O/gpu_tests.txt
