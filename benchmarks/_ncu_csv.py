"""print kernel name + duration (us) from an ncu --csv launch list"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split('(')[0][:70]
    key = (name, r[7], r[8])
    agg.setdefault(key, []).append(float(r[14]) / 1e3)
for (name, blk, grid), v in agg.items():
    v2 = sorted(v)
    print("%-72s blk=%s grid=%s n=%d med=%.2fus min=%.2f max=%.2f sum=%.1f" % (name, blk, grid, len(v), v2[len(v2)//2], v2[0], v2[-1], sum(v)))
if len(sys.argv) > 2:
    for r in rows[:int(sys.argv[2])]:
        print(r[0], r[4][:50], r[14])
