#!/usr/bin/env python
"""The few numbers of an `ncu --page raw --csv` dump that say what bounds a kernel."""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
d = dict(zip(rows[0], rows[2]))
for k in ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
          'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
          'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
          'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__warps_eligible.avg.per_cycle_active',
          'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
          'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'launch__registers_per_thread']:
    print(f"{k:75s} {d.get(k)}")
st = {k.split('stalled_')[1].split('_per')[0]: float(d[k]) for k in d if 'issue_stalled' in k and 'per_issue_active' in k}
tot = sum(st.values())
print('stalls per issue:', ', '.join(f"{k} {v:.2f}" for k, v in sorted(st.items(), key=lambda x: -x[1]) if v > 0.05))
