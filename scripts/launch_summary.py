"""Aggregate an ncu --metrics gpu__time_duration.sum CSV launch list by kernel name."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
h = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
H, data = rows[h], rows[h + 1:]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
agg = collections.OrderedDict()
for r in data:
    k = r[ki].split("(")[0]
    t = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)
    a = agg.setdefault(k, [0, 0.0, 0.0])
    a[0] += 1; a[1] += t; a[2] = max(a[2], t)
tot = sum(a[1] for a in agg.values())
print(f"launches {len(data)}  total {tot:.1f} us (ncu: cold cache, serialised -> compare shares)")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:28s} n={a[0]:4d} total={a[1]:9.1f} us share={100 * a[1] / tot:5.1f}% max={a[2]:8.1f} us")
