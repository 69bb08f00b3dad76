#!/bin/bash
# Round-2 evidence pass on the GPU box (profiles/README.md).  Outputs under gpurun_out/r3final/.
set -u
cd /root/repo
O=gpurun_out/r3final; mkdir -p $O
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
# launch list of the bench command (device time per launch, serialised, profiler cache flush off)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
# counters of the step kernel over bench.py's DRAM-cold graph replay (no profiler cache flush)
timeout 600 ncu --metrics sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none -k regex:gn_step_kernel -s 40 -c 32 --csv --log-file $O/gn_step_counts.csv python bench.py --steps 200 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
# one full capture of the step kernel inside the bench command
timeout 600 ncu --set full --import-source on --clock-control none --cache-control none -k regex:gn_step_kernel -s 60 -c 1 -o $O/gn_step python bench.py --steps 200 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
python scratch/ncu_summary.py $O/gn_step.ncu-rep $O/gn_step_full_summary.csv gn_step_kernel > /dev/null 2>&1
ncu -i $O/gn_step.ncu-rep --page details > $O/gn_step_details.txt 2>/dev/null
# K1 (hinge kernel) full capture
timeout 600 ncu --set full --clock-control none -k regex:hinge_kernel -s 3 -c 1 -o $O/hinge python scratch/r2_k1.py 32768 128 5 > $O/k1.log 2>&1
python scratch/ncu_summary.py $O/hinge.ncu-rep $O/hinge_full_summary.csv hinge_kernel > /dev/null 2>&1
# sanitizer over the round's new paths
{
echo '$ compute-sanitizer --tool memcheck: balanced waves, in-place host step, fused level 1'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_schedules.py tests/test_gpu_host_step.py -m gpu -q -k "balanced or in_place or chunked or fused" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | tail -5
echo '$ compute-sanitizer --tool racecheck: balanced waves, fused level 1'
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_schedules.py -m gpu -q -k "balanced or fused" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard|error" | tail -5
} > $O/sanitizer.txt 2>&1
rm -f $O/hinge.ncu-rep
ls -la $O; cat $O/sanitizer.txt; tail -c 600 $O/bench.err
