#!/usr/bin/env python
"""Summarise an .ncu-rep (full set) into a CSV + table: python tools/ncu_summary.py rep.ncu-rep out.csv"""
import csv, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
keep = ['ID', 'Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__waves_per_multiprocessor',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__issue_active.avg.pct_of_peak_sustained_active']
idx = [h.index(k) for k in keep if k in h]
with open(out, 'w') as f:
    w = csv.writer(f)
    w.writerow([h[i] for i in idx]); w.writerow([rows[1][i] for i in idx])
    for r in rows[2:]:
        w.writerow([r[i] for i in idx])
ki, ti = h.index('Kernel Name'), h.index('gpu__time_duration.sum')
unit = rows[1][ti]
tot = 0
for r in rows[2:]:
    v = float(r[ti].replace(',', ''))
    if unit == 'ms': v *= 1000
    elif unit == 'ns': v /= 1000
    tot += v
    print("%-3s %-44s %-16s %9.1f us  dramR %8s dramW %8s %s  sm%% %5s  warps%% %5s  thr/inst %5s" % (
        r[0], r[ki][:44], r[h.index('Grid Size')], v, r[h.index('dram__bytes_read.sum')][:8], r[h.index('dram__bytes_write.sum')][:8],
        rows[1][h.index('dram__bytes_read.sum')],
        r[h.index('sm__throughput.avg.pct_of_peak_sustained_elapsed')][:5], r[h.index('sm__warps_active.avg.pct_of_peak_sustained_active')][:5],
        r[h.index('smsp__thread_inst_executed_per_inst_executed.ratio')][:5]))
print("total %.1f us" % tot)
