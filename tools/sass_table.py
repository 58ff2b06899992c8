#!/usr/bin/env python
"""Static evidence from the built library, no GPU needed: per kernel of libh2agg.so the SASS instruction count, the counts
of the mnemonics that matter on this path, and the registers / stack / spills ptxas reported (halo2-aggregation_b200/build/*.log).

  python tools/sass_table.py > profiles/r2_sass_mnemonics.md

UBLKCP + SYNCS: 1-D TMA bulk copies completing on an mbarrier (cp.async.bulk ... mbarrier::complete_tx); IMAD.WIDE: the
32x32->64 multiply-adds of the Montgomery products (the pipe that binds); MATCH: warp-aggregated atomics; LDL / STL: local
memory (spills, or private arrays indexed at run time)."""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "halo2-aggregation_b200", "libh2agg.so")
KEYS = ["IMAD.WIDE", "IMAD", "IADD3", "UBLKCP", "SYNCS", "MATCH", "SHFL", "ATOMG", "LDG", "STG", "LDS", "STS", "BAR", "LDL", "STL"]


def short(name):
    name = name.replace("(anonymous namespace)::", "").replace("dev::", "")
    name = re.sub(r"^void ", "", name)
    depth, out = 0, []
    for ch in name:        # cut the argument list: the first '(' outside template brackets
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            break
        out.append(ch)
    return "".join(out)


def ptxas_info():
    info = {}
    for log in glob.glob(os.path.join(ROOT, "halo2-aggregation_b200", "build", "*.log")):
        text = open(log).read()
        for m in re.finditer(r"Compiling entry function '([^']+)' for 'sm_100a'\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                             r".*?Used (\d+) registers", text, re.S):
            info[m.group(1)] = tuple(int(m.group(i)) for i in (5, 2, 3, 4))
    return info


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    info = ptxas_info()
    parts = re.split(r"\n\s*Function : (\S+)\n", sass)
    rows = []
    for mangled, body in zip(parts[1::2], parts[2::2]):
        ops = re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", body, re.M)
        cnt = {k: 0 for k in KEYS}
        for op in ops:
            if op.startswith("IMAD.WIDE"):
                cnt["IMAD.WIDE"] += 1
            elif op.startswith("IMAD"):
                cnt["IMAD"] += 1
            else:
                for k in KEYS[2:]:
                    if op.startswith(k):
                        cnt[k] += 1
                        break
        regs = info.get(mangled, (None, None, None, None))
        rows.append((mangled, len(ops), cnt, regs))
    names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
    table = {}
    for (mangled, n, cnt, regs), dem in zip(rows, names):
        nm = short(dem)
        fam = re.sub(r"<\d+, ", "<C, ", nm)          # msm_hist_kernel<20, true> ... one row per family: the widest window
        if fam not in table or n > table[fam][1]:
            table[fam] = (nm, n, cnt, regs)
    print("# SASS per kernel of `libh2agg.so` (sm_100a): instruction and mnemonic counts, ptxas registers / stack / spills\n")
    print("`python tools/sass_table.py` (cuobjdump -sass + the `-Xptxas -v` logs of the build; no GPU needed). Templates over the MSM window width")
    print("are shown once (the largest instance). `UBLKCP` + `SYNCS` = 1-D TMA bulk copies completing on an mbarrier (the NTT's last pass); `IMAD.WIDE` = the")
    print("32x32->64 multiply-adds of the Montgomery products (the pipe that binds: 136 per product, 100 per square); `MATCH` = warp-aggregated atomics")
    print("(MSM histogram / scatter, lookup counting sort); `LDL` / `STL` = local memory (the division-step inversion's private arrays and XYZZ temporaries")
    print("indexed at run time; `spill` is what ptxas itself spilled).\n")
    print("| kernel | instr | regs | stack B | spill st/ld B | " + " | ".join(KEYS) + " |")
    print("|---|---|---|---|---|" + "---|" * len(KEYS))
    tot = {k: 0 for k in KEYS}
    for mangled, n, cnt, regs in rows:
        for k in KEYS:
            tot[k] += cnt[k]
    for fam, (nm, n, cnt, regs) in sorted(table.items(), key=lambda kv: -kv[1][1]):
        r = regs if regs[0] is not None else ("?", "?", "?", "?")
        print("| `%s` | %d | %s | %s | %s/%s | " % (nm, n, r[0], r[1], r[2], r[3]) + " | ".join(str(cnt[k]) for k in KEYS) + " |")
    tc = len(re.findall(r"\b(UTCMMA|UTCHMMA|HMMA|IMMA|QMMA|UTMALDG|UTMASTG)\b", sass))
    print("\nWhole library: %d kernels (%d template families), %s; tensor-core / tensor-map mnemonics (UTCMMA, HMMA, IMMA, UTMALDG, UTMASTG): %d — 254-bit modular"
          % (len(rows), len(table), ", ".join("%d `%s`" % (tot[k], k) for k in ("IMAD.WIDE", "UBLKCP", "SYNCS", "MATCH")), tc))
    print("arithmetic is not a dense contraction, and the only tile-shaped loads (NTT rows) are 1-D bulk copies.")


if __name__ == "__main__":
    sys.exit(main())
