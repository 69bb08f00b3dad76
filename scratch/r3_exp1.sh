#!/bin/bash
# round 2, session 2, call 1: balanced waves + early prefetch + L2 fetch granularity
mkdir -p gpurun_out/r3a; cd /root/repo
python -m pytest tests/test_gpu_schedules.py tests/test_gpu_parity.py -m gpu -x -q -s 2>&1 | tail -15 > gpurun_out/r3a/pytest.log
python scratch/r3_cold.py 64 1024 DGPMP2_PREFETCH=3 DGPMP2_PREFETCH=2 > gpurun_out/r3a/cold_T64.json 2>gpurun_out/r3a/cold_T64.err
python scratch/r3_cold.py 64 1024 gran=32 DGPMP2_PREFETCH=3 > gpurun_out/r3a/cold_T64_g32.json 2>>gpurun_out/r3a/cold_T64.err
python scratch/r3_cold.py 64 1024 gran=128 > gpurun_out/r3a/cold_T64_g128.json 2>>gpurun_out/r3a/cold_T64.err
python scratch/r3_cold.py 128 1024 DGPMP2_BALANCED=2 DGPMP2_PREFETCH=3 > gpurun_out/r3a/cold_T128.json 2>gpurun_out/r3a/cold_T128.err
cat gpurun_out/r3a/pytest.log gpurun_out/r3a/*.json; tail -3 gpurun_out/r3a/*.err
