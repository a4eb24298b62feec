"""Join an ncu SASS source page (csv) with nvdisasm -g line info: executed warp instructions and stall samples per
source line.  usage: _ncu_lines.py <src.csv> <nvdisasm.txt> <mangled-name> [top]"""
import csv, re, sys, collections
src, dis, fn = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
line_of = {}
cur = None; infn = False
for l in open(dis):
    if l.startswith('//----') and '.text.' in l:
        infn = ('.text.' + fn + ' ') in l or l.strip().endswith(fn + ' --------------------------') or (fn in l)
        continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', l)
    if m: line_of[int(m.group(1), 16)] = cur
rr = list(csv.reader(open(src)))
h = rr[1]; body = []
for r in rr[2:]:
    if r and r[0] == 'Kernel Name': break
    body.append(r)
ia, ie, iss = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples")
base = int(body[0][ia], 16)
agg = collections.defaultdict(lambda: [0, 0, 0])
for r in body:
    off = int(r[ia], 16) - base
    k = line_of.get(off)
    agg[k][0] += int(r[ie]); agg[k][1] += int(r[iss]); agg[k][2] += 1
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print("total warp instr %d samples %d" % (ti, ts))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-28s inst %5.1f%%  samples %5.1f%%  sass %d" % (k, 100 * v[0] / ti, 100 * v[1] / ts, v[2]))
