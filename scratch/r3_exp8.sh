#!/bin/bash
mkdir -p gpurun_out/r3t; cd /root/repo; O=gpurun_out/r3t
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/pytest.log
python scratch/r3_bits.py dump /tmp/new.pt | tail -1; DGPMP2_LIB=scratch/exp/lib_prev.so python scratch/r3_bits.py dump /tmp/prev.pt | tail -1
python scratch/r3_bits.py cmp /tmp/prev.pt /tmp/new.pt > $O/bits.txt
python scratch/r3_cold.py 64 1024 > $O/cold.json; DGPMP2_LIB=scratch/exp/lib_prev.so python scratch/r3_cold.py 64 1024 > $O/cold_prev.json
python scratch/r3_cfg.py config3_point_T128 > $O/cfg3.json; DGPMP2_LIB=scratch/exp/lib_prev.so python scratch/r3_cfg.py config3_point_T128 > $O/cfg3_prev.json
python scratch/r3_cfg.py config4_nonholonomic_T96 > $O/cfg4.json; DGPMP2_LIB=scratch/exp/lib_prev.so python scratch/r3_cfg.py config4_nonholonomic_T96 > $O/cfg4_prev.json
timeout 300 ncu --metrics l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,gpu__time_duration.sum --clock-control none -k regex:gn_step_kernel -s 10 -c 2 --csv --log-file $O/conflicts.csv python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
cat $O/pytest.log $O/bits.txt $O/cold.json $O/cold_prev.json; for f in cfg3 cfg3_prev cfg4 cfg4_prev; do python -c "
import json; d=json.load(open('$O/$f.json')); print('$f', d['us_per_step']['default'])"; done
grep -v "^==" $O/conflicts.csv | cut -d, -f13- | tail -6
