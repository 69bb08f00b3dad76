#!/bin/bash
set -u
O=gpurun_out/call3; mkdir -p $O
timeout 120 python -m pytest tests/test_gpu_api.py -m gpu -q -x -k "sdf_generation or headless" 2>&1 | tail -3 | tee $O/edt_tests.txt
DGPMP2_LIB=$PWD/scratch/lib_base.so timeout 100 python scratch/edt_time.py 2>&1 | sed 's/^/old: /' | tee $O/edt_time.txt
timeout 100 python scratch/edt_time.py 2>&1 | sed 's/^/new: /' | tee -a $O/edt_time.txt
