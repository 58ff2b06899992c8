#!/usr/bin/env python
"""Aggregates an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel name.
usage: launch_shares.py list.csv [first_index [last_index]]
       launch_shares.py list.csv --proof      (the launches of ONE proof of tools/prove_bench.py: everything between
                                               the last two quotient kernels, i.e. one period of the prover's sequence)"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 8]
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0}
names = [(r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "").replace("dev::", ""),
          float(r[vi].replace(",", "")) * scale.get(r[ui], 1e-6)) for r in rows[1:]]
groups = None
if len(sys.argv) > 2 and sys.argv[2] == "--proof":
    q = [i for i, (n, _) in enumerate(names) if n.startswith("quotient_kernel")]
    if len(q) < 2:
        sys.exit("need two proofs in the list")
    lo, hi = q[-2], q[-1]

    def group(n):
        if n.startswith(("msm_", "aff_", "scan_block", "scan_top", "scan_apply")):
            return "MSM (commitments)"
        if n.startswith("ntt_"):
            return "NTT (ifft + coset fft)"
        if n.startswith("quotient"):
            return "quotient"
        if n.startswith(("bitonic", "count_", "expand_", "key_", "lookup_", "u32_", "fill_u32", "compress")):
            return "lookup permutation (sort, match)"
        return "other (grand products, evaluations, Kate division)"
    groups = collections.OrderedDict()
else:
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else len(names)
    if lo < 0:
        lo += len(names)
agg = collections.OrderedDict()
for n, t in names[lo:hi]:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += t
    if groups is not None:
        groups[group(n)] = groups.get(group(n), 0.0) + t
tot = sum(a[1] for a in agg.values())
print("launches %d..%d of %d, total %.2f ms" % (lo, hi, len(names), tot))
if groups is not None:
    print("| group | total ms | share |\n|---|---|---|")
    for g, t in sorted(groups.items(), key=lambda kv: -kv[1]):
        print("| %s | %.2f | %.1f%% |" % (g, t, 100 * t / tot))
    print()
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| %s | %d | %.3f | %.1f%% |" % (n, c, t, 100 * t / tot))
