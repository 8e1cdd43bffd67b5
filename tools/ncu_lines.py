"""Instructions and stall samples per SOURCE LINE of one kernel of an .ncu-rep (captured with --import-source on, library built
with -lineinfo): the SASS listing of the report's source page is matched, instruction by instruction, with `nvdisasm -g` of the
kernel's cubin extracted from the in-tree library (ncu's CSV export of the CUDA view carries no metrics).

    python tools/ncu_lines.py <report.ncu-rep> <source file in csrc, e.g. knn.cu> <substring of the mangled kernel name> [top N]
"""
import collections, csv, os, re, subprocess, sys, tempfile

rep, srcname, kern = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, 'synthobs_b200', 'libsynthobs_b200.so')
tmp = tempfile.mkdtemp()
stem = srcname[:-3]
subprocess.run('cd %s && cuobjdump -xelf all %s > /dev/null 2>&1 && nvdisasm -g -c %s.sm_100a.cubin > k.dis 2> /dev/null' % (tmp, lib, stem),
               shell=True, check=True)
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, v = rows[0], rows[2]
print('kernel', v[hdr.index('Kernel Name')][:100])
for m in ('gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
          'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
          'launch__registers_per_thread'):
    print('  %-62s %s' % (m, v[hdr.index(m)]))
page = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(page.splitlines()))
hdr = rows[1]
ii, ss = hdr.index('Instructions Executed'), hdr.index('# Samples')
data = rows[2:]
dis = open(os.path.join(tmp, 'k.dis')).read().split('\n')
start = [i for i, l in enumerate(dis) if l.startswith('.text.') and kern in l][0]
cur, seq = None, []
for l in dis[start + 1:]:
    if (l.startswith('.text.') or l.startswith('//--------------------- .text')) and seq:
        break
    m = re.search(r'//## File ".*%s", line (\d+)' % re.escape(srcname), l)
    if m:
        cur = int(m.group(1))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+.*?;', l):
        seq.append(cur)
assert len(seq) == len(data), (len(seq), len(data))
agg = collections.defaultdict(lambda: [0, 0])
for line, r in zip(seq, data):
    agg[line][0] += int(r[ii])
    agg[line][1] += int(r[ss])
tot = sum(a[0] for a in agg.values())
stot = sum(a[1] for a in agg.values())
src = open(os.path.join(root, 'synthobs_b200', 'csrc', srcname)).read().split('\n')
print('| line | warp instructions | share | stall samples | source |')
print('|---|---|---|---|---|')
for line, (a, b) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print('| %s | %d | %.1f %% | %.1f %% | `%s` |' % (line, a, 100 * a / tot, 100 * b / stot, src[line - 1].strip()[:100].replace('|', '¦') if line else ''))
