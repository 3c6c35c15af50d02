#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` dump: the handful of metrics the roofline needs."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
H = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__waves_per_multiprocessor', 'smsp__inst_executed.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum']
ki = H.index('Kernel Name')
stalls = [(i, h) for i, h in enumerate(H) if 'issue_stalled' in h and h.endswith('_ratio') is False and 'pct' in h]
if not stalls:
    stalls = [(i, h) for i, h in enumerate(H) if 'warp_issue_stalled' in h]
for r in rows[2:]:
    print('==', r[ki].split('(')[0], 'units', )
    for w in want:
        if w in H:
            print('   %-70s %s %s' % (w, r[H.index(w)], rows[1][H.index(w)]))
    vals = []
    for i, h in stalls:
        try:
            vals.append((float(r[i].replace(',', '')), h))
        except ValueError:
            pass
    for v, h in sorted(vals, reverse=True)[:6]:
        print('   stall %-64s %.2f' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('smsp__average_warp_latency_issue_stalled_', ''), v))
