"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python tools/launch_table.py file.csv [top]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
n = rows[h]
k, mv, gs = n.index("Kernel Name"), n.index("Metric Value"), n.index("Grid Size")
agg, tot, cnt, seq = collections.OrderedDict(), 0.0, 0, []
for r in rows[h + 1:]:
    if len(r) <= mv:
        continue
    name = re.sub(r"\(.*", "", r[k]).replace("void ", "").replace("unnamed>::", "")
    v = float(r[mv].replace(",", "")) / 1000
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
    cnt += 1
    seq.append((name, v, r[gs]))
print(f"launches {cnt}  sum of kernel times {tot:.1f} us")
for nm, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 12]:
    print(f"{nm:42s} n={c:3d} {t / c:7.1f} us/launch {t:8.1f} us  {100 * t / tot:5.1f}%")
if len(sys.argv) > 3:
    for nm, v, g in seq[: int(sys.argv[3])]:
        print(f"   {nm:40s} {v:7.1f} us  grid {g}")
