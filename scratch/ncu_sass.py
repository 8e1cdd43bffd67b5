import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ie = hdr.index('Instructions Executed'); si = hdr.index('Source'); ss = hdr.index('# Samples')
tot = sum(float(r[ie]) for r in rows[2:] if len(r) > ie and r[ie].replace('.','').isdigit())
top = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
byop = collections.Counter()
print('total warp-instr', tot)
for k, r in enumerate(rows[2:]):
    if len(r) <= ie: continue
    try: v = float(r[ie])
    except: continue
    op = r[si].split()[0] if not r[si].strip().startswith('@') else r[si].split()[1]
    byop[op.split('.')[0]] += v
    if v / tot * 100 >= top:
        print('%4d %5.2f%% smp=%5s %s' % (k, v / tot * 100, r[ss], r[si].strip()[:90]))
print('--- by opcode')
for op, v in byop.most_common(22):
    print('%-10s %5.1f%%' % (op, 100 * v / tot))
