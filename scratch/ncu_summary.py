"""Compact, committed summary of an ncu report: python scratch/ncu_summary.py in.ncu-rep out.csv [kernel-regex]"""
import csv, re, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
pat = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u = rows[0], rows[1]
KEEP = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__average_warp_latency_issue_stalled_barrier.pct',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max',
        'smsp__inst_executed_per_warp.ratio', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_lsu.sum',
        'smsp__warp_issue_stalled_barrier_per_warp_active.pct', 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_wait_per_warp_active.pct',
        'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct', 'local_load_bytes', 'local_store_bytes',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']
idx = [i for i, n in enumerate(h) if n in KEEP]
with open(out, 'w') as f:
    w = csv.writer(f)
    w.writerow(['metric', 'unit'] + ['launch %d' % k for k in range(len(rows) - 2)])
    sel = [r for r in rows[2:] if pat is None or pat.search(r[h.index('Kernel Name')])]
    for i in idx:
        w.writerow([h[i], u[i]] + [r[i] for r in sel])
print(open(out).read())
