#!/usr/bin/env python
"""Builds a kernel variant of libh2agg.so for A/B measurement: tools/build_variant.py NAME -DTREE_PF_ROUND=0 ...
-> tools/variants/libh2agg_NAME.so (git-ignored, travels with gpurun); run with H2A_LIB=tools/variants/libh2agg_NAME.so."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "halo2-aggregation_b200"))
import _build as b

name, extra = sys.argv[1], sys.argv[2:]
csrc = b.CSRC
if extra and extra[0].startswith("--csrc="):     # another source tree (e.g. a checkout of an earlier commit's csrc/)
    csrc, extra = extra[0][len("--csrc="):], extra[1:]
out_dir = os.path.join(ROOT, "tools", "variants")
obj_dir = os.path.join(out_dir, "obj_" + name)
os.makedirs(obj_dir, exist_ok=True)


def one(src):
    obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
    flags = [f for f in b.FLAGS if f not in ("-Xptxas", "-v")]
    r = subprocess.run([b.NVCC] + flags + extra + ["-I", os.path.join(ROOT, "include"), "-c", os.path.join(csrc, src), "-o", obj], capture_output=True, text=True)
    if r.returncode:
        raise SystemExit(r.stderr[-3000:])
    return obj


with ThreadPoolExecutor(8) as ex:
    objs = list(ex.map(one, b.SOURCES))
out = os.path.join(out_dir, "libh2agg_%s.so" % name)
subprocess.check_call([b.NVCC, "-shared", "-o", out] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-ldl"])
print(out)
