#!/usr/bin/env python
"""Print the handful of ncu metrics DESIGN.md quotes from a .ncu-rep: python tools/ncu_brief.py file.ncu-rep [more ...]"""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'dram__throughput.avg.pct_of_peak_sustained_elapsed']
for f in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', f, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    for row in rows[2:]:
        print('==', f, row[h.index('Kernel Name')][:60])
        for k in KEYS:
            if k in h: print('   %-75s %s %s' % (k, row[h.index(k)], u[h.index(k)]))
        st = [(float(row[i]), n.split('issue_stalled_')[1].split('_per')[0]) for i, n in enumerate(h)
              if n.startswith('smsp__average_warps_issue_stalled_') and n.endswith('_per_issue_active.ratio') and 'not_issued' not in n and row[i]]
        print('   stalls/issue:', ', '.join('%s %.2f' % (n, v) for v, n in sorted(st, reverse=True)[:7]))
