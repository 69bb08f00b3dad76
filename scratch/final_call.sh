#!/bin/bash
# Round-end evidence pass for the shipped build (bounded: every step has its own timeout).  Outputs: gpurun_out/final2/.
set -u
O=gpurun_out/final2; mkdir -p $O
t0=$(date +%s); el() { echo $(( $(date +%s)-t0 )); }
timeout 100 python -m pytest tests -m gpu -q > $O/gpu_tests.txt 2>&1; echo "gpu tests rc=$? $(el)s" | tee -a $O/log.txt; tail -3 $O/gpu_tests.txt
timeout 150 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$? $(el)s" | tee -a $O/log.txt; cat $O/bench.json | cut -c1-600
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1; echo "ncu launches rc=$? $(el)s" | tee -a $O/log.txt
timeout 60 ncu --metrics sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gn_step_kernel -s 8 -c 1 --csv --log-file $O/gn_step_counts.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo "ncu counts rc=$? $(el)s" | tee -a $O/log.txt
if [ $(el) -lt 300 ]; then timeout 50 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$? $(el)s" | tee -a $O/log.txt; fi
if [ $(el) -lt 290 ]; then timeout 45 python scratch/extra_timings.py > $O/extra_timings.json 2> $O/extra_timings.err; echo "extra rc=$? $(el)s" | tee -a $O/log.txt; fi
if [ $(el) -lt 260 ]; then timeout 90 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_head.py tests/test_gpu_launch.py -m gpu -q -k "golden or bitwise or eager" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid" | tail -4 > $O/sanitizer_head.txt; echo "sanitizer $(el)s" | tee -a $O/log.txt; cat $O/sanitizer_head.txt; fi
ls $O
