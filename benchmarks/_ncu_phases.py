"""Executed warp instructions of an ncu SASS source page by address-ordered source-line region: prints, in address order,
runs of SASS instructions with their source line, executed count and opcode mix, merged into chunks.
usage: _ncu_phases.py <src.csv> <nvdisasm.txt> <mangled-name> [threads_total]"""
import csv, sys, collections, re
src, dis, fn = sys.argv[1:4]
line_of = {}; cur = None; infn = False
for l in open(dis):
    if l.startswith('//----') and '.text.' in l:
        infn = fn in l; continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', l)
    if m: line_of[int(m.group(1), 16)] = cur
rr = list(csv.reader(open(src))); h = rr[1]
body = [r for r in rr[2:] if r and r[0].startswith('0x')]
ia, ie, isrc, iss = h.index("Address"), h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
base = int(body[0][ia], 16)
tot = sum(int(r[ie]) for r in body)
warps = int(body[0][ie])          # first instruction: executed once per warp
print("warps %d, warp-instr %d, per warp %.0f" % (warps, tot, tot / warps))
# chunk by executed count (a change of count = a different loop depth / branch region)
chunks = []
for r in body:
    off = int(r[ia], 16) - base; e = int(r[ie]); op = r[isrc].split()[0] if not r[isrc].strip().startswith('@') else r[isrc].split()[1]
    op = op.split('.')[0]
    ln = line_of.get(off)
    if chunks and chunks[-1]['e'] == e:
        c = chunks[-1]
    else:
        c = {'e': e, 'n': 0, 'ops': collections.Counter(), 'lines': collections.Counter(), 'start': off, 's': 0}; chunks.append(c)
    c['n'] += 1; c['ops'][op] += 1; c['s'] += int(r[iss])
    if ln: c['lines'][ln[0][:10] + ':' + str(ln[1])] += 1
for c in chunks:
    if c['n'] * c['e'] < 0.002 * tot: continue
    print("@%05x n=%4d exec/warp=%6.2f  share=%5.1f%% samp=%5d  ops: %s | %s" % (c['start'], c['n'], c['e'] / warps, 100.0 * c['n'] * c['e'] / tot, c['s'],
          " ".join("%s%d" % kv for kv in c['ops'].most_common(8)), " ".join("%s" % k for k, _ in c['lines'].most_common(4))))
