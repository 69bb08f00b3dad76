#!/bin/bash
# round 2, session 2, call 2: defaults = early prefetch + balanced waves; d=6 / T=128 knob sweeps; full GPU suite
mkdir -p gpurun_out/r3b; cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r3b/pytest.log
python scratch/r3_cfg.py config4_nonholonomic_T96 DGPMP2_WIDE=100 DGPMP2_WIDE=200 DGPMP2_THREADS=256 DGPMP2_WIDE=100,DGPMP2_THREADS=256 DGPMP2_TAIL=3 DGPMP2_TAIL=6 DGPMP2_NP=1 DGPMP2_NP=1,DGPMP2_THREADS=128 > gpurun_out/r3b/cfg4.json 2>gpurun_out/r3b/cfg4.err
python scratch/r3_cfg.py config3_point_T128 DGPMP2_WIDE=100 DGPMP2_WIDE=28 DGPMP2_TAIL=2 DGPMP2_TAIL=8 DGPMP2_THREADS=384 > gpurun_out/r3b/cfg3.json 2>gpurun_out/r3b/cfg3.err
python scratch/r3_cold.py 64 1024 DGPMP2_PDL=2 DGPMP2_PREFETCH=3 > gpurun_out/r3b/cold_T64.json 2>gpurun_out/r3b/cold.err
cat gpurun_out/r3b/pytest.log gpurun_out/r3b/*.json; tail -n 3 gpurun_out/r3b/*.err
