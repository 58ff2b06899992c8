#!/usr/bin/env python
"""Aggregates an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel name.
usage: launch_shares.py list.csv [first_index [last_index]]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 8]
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0}
names = [(r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", ""), float(r[vi].replace(",", "")) * scale.get(r[ui], 1e-6))
         for r in rows[1:]]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else len(names)
if lo < 0:
    lo += len(names)
agg = collections.OrderedDict()
for n, t in names[lo:hi]:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(a[1] for a in agg.values())
print("launches %d..%d of %d, total %.2f ms" % (lo, hi, len(names), tot))
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| %s | %d | %.3f | %.1f%% |" % (n, c, t, 100 * t / tot))
