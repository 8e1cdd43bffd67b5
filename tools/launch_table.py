"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list as a markdown table."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1], errors='replace')))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]
kn, mv = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    try:
        v = float(r[mv].replace(',', ''))
    except ValueError:
        continue
    name = r[kn].replace('<unnamed>::', '').split('(')[0][:60]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
n = sum(a[0] for a in agg.values())
print('%s\n\n%d launches, %.1f ms of kernel time in total (setup, warm-up and timed steps of all workloads; cold-cache, serialised: compare shares).\n' % (sys.argv[2] if len(sys.argv) > 2 else '', n, tot / 1e6))
print('| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|')
for name, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print('| %s | %d | %.3f | %.1f | %.1f%% |' % (name, c, v / 1e6, v / c / 1e3, 100 * v / tot))
