"""Summarises an .ncu-rep (raw page) per launch: python profiles/ncu_summary.py file.ncu-rep [launch ids...]"""
import csv, subprocess, sys
rep = sys.argv[1]
ids = [int(x) for x in sys.argv[2:]]
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
H = rows[0]
want = ['Kernel Name', 'launch__grid_size', 'launch__registers_per_thread', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']
for n, r in enumerate(rows[2:]):
    if ids and n not in ids:
        continue
    print('--- launch', n)
    for w in want:
        if w in H:
            print('  %-66s %s %s' % (w, r[H.index(w)][:70], rows[1][H.index(w)]))
    vals = []
    for i, h in enumerate(H):
        if 'smsp__pcsamp_warps_issue_stalled' in h and not h.endswith('_not_issued'):
            try:
                vals.append((float(r[i].replace(',', '')), h))
            except ValueError:
                pass
    tot = sum(v for v, _ in vals) or 1
    print('  stalls: ' + ', '.join('%s %.0f%%' % (h.replace('smsp__pcsamp_warps_issue_stalled_', ''), 100 * v / tot) for v, h in sorted(vals, reverse=True)[:8]))
