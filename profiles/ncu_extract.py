#!/usr/bin/env python
"""Pull the handful of counters we track out of an `ncu --page raw --csv` export.
usage: python profiles/ncu_extract.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
H = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum',
        'smsp__warps_eligible.avg.per_cycle_active', 'lts__t_bytes.sum']
want += [h for h in H if 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct')]
idx = {h: i for i, h in enumerate(H)}
for w in want:
    if w in idx:
        print("%-78s %s [%s]" % (w[:78], [r[idx[w]][:16] for r in rows[2:]], rows[1][idx[w]]))
    else:
        print("MISSING", w)
