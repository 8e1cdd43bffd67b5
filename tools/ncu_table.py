"""Summarise an .ncu-rep (first launch of every distinct kernel) as a markdown table."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = open(rep, errors='replace').read() if rep.endswith('.csv') else \
    subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
cols = [('gpu__time_duration.sum', 'time'), ('dram__bytes_read.sum', 'dram rd'), ('dram__bytes_write.sum', 'dram wr'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram %'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm %'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue %'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor %'),
        ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'xu %'),
        ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'fma %'),
        ('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'fp64 %'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ %'),
        ('launch__registers_per_thread', 'regs'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
        ('smsp__inst_executed.sum', 'warp inst')]
seen = {}
for r in rows[2:]:
    name = r[idx['Kernel Name']]
    short = name.replace('<unnamed>::', '').split('(')[0][:44]
    seen.setdefault(short, r)     # keep the LAST? first is cold; prefer last launch
    seen[short] = r
print('| kernel | ' + ' | '.join(c[1] for c in cols) + ' |')
print('|---|' + '---|' * len(cols))
for short, r in seen.items():
    vals = []
    for m, _ in cols:
        if m in idx:
            v = r[idx[m]]; u = units[idx[m]]
            try:
                f = float(v.replace(',', ''))
                v = ('%.3g' % f) if abs(f) < 1e6 else ('%.3e' % f)
            except ValueError:
                pass
            vals.append(v + (' ' + u if u and u not in ('%', 'inst') and not u.startswith('register') and u not in ('block',) else ''))
        else:
            vals.append('-')
    print('| %s | ' % short + ' | '.join(vals) + ' |')
