#!/usr/bin/env python
"""Pivot an `ncu --csv --metrics ...` log into one row per launch: usage ncu_table.py log.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 8]
h = rows[0]
idx = {k: h.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "Grid Size")}
per = collections.OrderedDict()
for r in rows[1:]:
    key = int(r[idx["ID"]])
    d = per.setdefault(key, {"name": r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", ""), "grid": r[idx["Grid Size"]]})
    d[r[idx["Metric Name"]]] = (r[idx["Metric Value"]].replace(",", ""), r[idx["Metric Unit"]])
short = [("gpu__time_duration.sum", "us"), ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "fmaH%"),
         ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("dram__bytes_read.sum", "rdGB"), ("dram__bytes_write.sum", "wrGB"),
         ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
         ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_lsb"),
         ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
         ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"), ("lts__t_sector_hit_rate.pct", "L2hit%"),
         ("l1tex__t_sector_hit_rate.pct", "L1hit%"), ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "Minst")]
print("| kernel | grid | " + " | ".join(s for _, s in short) + " |")
print("|---|---|" + "---|" * len(short))
tot = 0.0
for d in per.values():
    cells = []
    for m, s in short:
        v, u = d.get(m, ("", ""))
        try:
            f = float(v)
            if s == "us":
                f *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0); tot += f
            if s in ("rdGB", "wrGB"):
                f *= {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(u, 1e-9)
            if s == "Minst":
                f *= 1e-6
            cells.append("%.2f" % f if s != "regs" else "%d" % f)
        except ValueError:
            cells.append(v)
    print("| %s | %s | %s |" % (d["name"], d["grid"], " | ".join(cells)))
print("total %.1f us" % tot)
