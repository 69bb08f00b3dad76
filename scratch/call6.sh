#!/bin/bash
set -u
O=gpurun_out/final3; mkdir -p $O
t0=$(date +%s); el() { echo $(( $(date +%s)-t0 )); }
timeout 40 python -m pytest tests/test_gpu_launch.py -m gpu -q 2>&1 | tail -1 | tee $O/launch_tests.txt
timeout 60 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$? $(el)s" | tee -a $O/log.txt; cut -c1-200 $O/bench.json
timeout 45 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1; echo "ncu launches rc=$? $(el)s" | tee -a $O/log.txt
if [ $(el) -lt 80 ]; then timeout 40 ncu --metrics sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gn_step_kernel -s 8 -c 1 --csv --log-file $O/gn_step_counts.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo "ncu counts rc=$? $(el)s" | tee -a $O/log.txt; fi
