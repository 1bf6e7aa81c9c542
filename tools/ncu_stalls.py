#!/usr/bin/env python
"""Per-kernel duration, throughput and top stall reasons from `ncu -i rep --page raw --csv` output.
Usage: python tools/ncu_stalls.py raw.csv [kernel substring ...]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
pats = sys.argv[2:]
h = rows[0]
want = [c for c in h if 'smsp__average_warp' in c and 'issue_stalled' in c and c.endswith('.ratio') and 'not_issued' not in c]
extra = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
         'smsp__issue_active.avg.per_cycle_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
         'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
         'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
         'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__grid_size', 'launch__block_size']
ki = h.index('Kernel Name')
for r in rows[2:]:
    name = r[ki][:50]
    if pats and not any(p in name for p in pats):
        continue
    print('==', r[0], name)
    st = sorted(((float(r[h.index(c)].replace(',', '')), c.split('issue_stalled_')[1].replace('_per_warp_active.ratio', ''))
                 for c in want), reverse=True)[:7]
    print('   stalls:', ', '.join('%s=%.2f' % (n, v) for v, n in st))
    print('   ' + ', '.join("%s=%s%s" % (e.split('__')[1][:34], r[h.index(e)][:12], rows[1][h.index(e)]) for e in extra if e in h))
