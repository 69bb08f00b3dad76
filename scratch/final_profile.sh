#!/bin/bash
# Round-end evidence pass on the GPU box (see profiles/README.md).  Outputs under gpurun_out/final/.
set -u
O=gpurun_out/final; mkdir -p $O
K="learned or oob or T3 or T101 or T128 or qfull or reference_autograd or nonholonomic or vel_limits or forward_batch or alignment"
{
echo '$ compute-sanitizer --tool memcheck (default schedule)'
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_backward.py -m gpu -q -k "$K" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | tail -5
echo '$ DGPMP2_WIDE=4 DGPMP2_NP=3 compute-sanitizer --tool memcheck (one-lane levels and multi-problem CTAs forced on the small cases)'
DGPMP2_WIDE=4 DGPMP2_NP=3 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_backward.py -m gpu -q -k "$K" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | tail -5
echo '$ compute-sanitizer --tool racecheck (default schedule)'
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_backward.py -m gpu -q -k "learned or oob or T3 or T101 or T128 or reference_autograd or forward_batch" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard|error" | tail -5
echo '$ DGPMP2_WIDE=4 DGPMP2_NP=3 compute-sanitizer --tool racecheck'
DGPMP2_WIDE=4 DGPMP2_NP=3 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_backward.py -m gpu -q -k "learned or oob or T3 or T101 or T128 or reference_autograd or forward_batch" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard|error" | tail -5
} > $O/sanitizer.txt 2>&1
# launch list of the bench command
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
# one full capture of the hot kernel inside the bench command
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gn_step_kernel -s 8 -c 1 -o $O/gn_step python bench.py --steps 20 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --metrics sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gn_step_kernel -s 8 -c 1 --csv --log-file $O/gn_step_counts.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
# micro-benchmarks behind the latency / issue model
{ echo '$ scratch/ubench5 (DFMA, shared operands)'; timeout 60 ./scratch/ubench5; echo; echo '$ scratch/ubench7 (fp64 issue rate by operand pattern)'; timeout 60 ./scratch/ubench7; echo; echo '$ scratch/ubench6 (block-Thomas forward step, primitive chains)'; timeout 60 ./scratch/ubench6; } > $O/microbench.txt 2>&1
timeout 300 python scratch/extra_timings.py > $O/extra_timings.json 2> $O/extra_timings.err
for i in 1 2 3; do timeout 60 python scratch/phase_time.py 1024 64; done > $O/phase_1024.txt 2>&1
for i in 1 2 3; do timeout 60 python scratch/phase_time.py 1 64; done > $O/phase_1.txt 2>&1
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
ls -la $O
