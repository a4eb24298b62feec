"""Stall-reason totals of an ncu SASS source page (csv), overall and for the SASS lines of selected source lines.
usage: _ncu_stalls.py <src.csv> [<nvdisasm.txt> <mangled-name>]"""
import csv, sys, collections, re
rr = list(csv.reader(open(sys.argv[1])))
h = rr[1]
body = [r for r in rr[2:] if r and r[0].startswith('0x')]
cols = [i for i, n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n]
tot = collections.Counter()
for r in body:
    for i in cols:
        tot[h[i]] += int(r[i] or 0)
s = sum(tot.values())
print("samples", s)
for k, v in tot.most_common():
    print("  %-24s %5.1f%%" % (k, 100.0 * v / s))
if len(sys.argv) > 3:
    dis, fn = sys.argv[2:4]
    line_of = {}; cur = None; infn = False
    for l in open(dis):
        if l.startswith('//----') and '.text.' in l:
            infn = fn in l; continue
        if not infn: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', l)
        if m: line_of[int(m.group(1), 16)] = cur
    ia = h.index("Address"); base = int(body[0][ia], 16)
    per = collections.defaultdict(collections.Counter)
    for r in body:
        k = line_of.get(int(r[ia], 16) - base)
        for i in cols: per[k][h[i]] += int(r[i] or 0)
    for k, c in sorted(per.items(), key=lambda kv: -sum(kv[1].values()))[:25]:
        t = sum(c.values())
        print("%-26s %5.1f%%  " % (k, 100.0 * t / s) + " ".join("%s=%.0f%%" % (n[6:], 100.0 * v / t) for n, v in c.most_common(4)))
