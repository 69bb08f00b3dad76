"""Attribute an ncu --import-source capture to source lines.
usage: python scratch/ncu_lines.py <sass.csv from `ncu -i rep --page source --csv --print-source sass`>
                                   <nvdisasm --print-line-info dump> <mangled kernel name prefix>"""
import re, csv, collections, sys
sass_csv, dis, kname = sys.argv[1], sys.argv[2], sys.argv[3]
lines = open(dis).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('.text.' + kname))
cur = None; stack = []; seq = {}
for l in lines[start + 1:]:
    if l.startswith('//---------------------'): break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)), m.group(3)); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: seq[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]; iA = hdr.index('Address'); iI = hdr.index('Instructions Executed'); iS = hdr.index('# Samples')
base = int(rows[2][iA], 16)
per = collections.Counter(); pers = collections.Counter(); ops = collections.Counter(); tot = 0
for r in rows[2:]:
    off = int(r[iA], 16) - base; n = int(r[iI]); s = int(r[iS]); tot += n
    c, txt = seq.get(off, (None, '?'))
    key = (c[0], c[1]) if c else ('?', 0)
    per[key] += n; pers[key] += s
    op = txt.split()[1] if txt.startswith('@') else txt.split()[0]
    ops[op.split('.')[0]] += n
src = {}
for f in ('bcr.cuh', 'kernels.cuh', 'factors.cuh', 'bcr_plan.cuh', 'hd.cuh', 'mp.cuh'):
    src[f] = open('/root/repo/dgpmp2_b200/csrc/' + f).read().split('\n')
print('total warp instructions', tot)
print('opcode mix:', ', '.join('%s %.1f%%' % (k, 100 * v / tot) for k, v in ops.most_common(22)))
ts = sum(pers.values())
for (f, ln), n in per.most_common(int(sys.argv[4]) if len(sys.argv) > 4 else 45):
    s = src[f][ln - 1].strip()[:90] if f in src else ''
    print('%-12s %4d %7d %4.1f%% smp %4.1f%% | %s' % (f, ln, n, 100 * n / tot, 100 * pers[(f, ln)] / ts, s))
