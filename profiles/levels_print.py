import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 8]
H = rows[0]; ki = H.index('Kernel Name'); mi = H.index('Metric Name'); vi = H.index('Metric Value'); ii = H.index('ID'); gi = H.index('Grid Size')
d = {}
for r in rows[1:]:
    d.setdefault((int(r[ii]), r[ki][:34], r[gi]), {})[r[mi]] = float(r[vi].replace(',', ''))
tot = 0
for k, v in sorted(d.items()):
    t = v['gpu__time_duration.sum'] / 1e3; rd = v['dram__bytes_read.sum']; wr = v['dram__bytes_write.sum']
    tot += t
    print('%-3d %-34s %-14s us=%8.1f GB=%6.2f TB/s=%5.2f inst=%.3g issue=%.1f%% l1=%.1f%% warps=%.1f%%' % (
        k[0], k[1], k[2], t, (rd + wr) / 1e9, (rd + wr) / max(t, 1e-9) / 1e6, v['smsp__inst_executed.sum'],
        v['smsp__issue_active.avg.pct_of_peak_sustained_active'], v['l1tex__throughput.avg.pct_of_peak_sustained_elapsed'],
        v['sm__warps_active.avg.pct_of_peak_sustained_active']))
print('total us', tot)
