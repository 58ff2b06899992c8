#!/usr/bin/env python
"""Short fixed workload for ncu: 2 MSMs of 2^22 points and 2 NTTs of 2^22 elements on device-resident data."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import halo2_aggregation_b200 as h2a

ctx = h2a.Context(0)
lg = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 22
n = 1 << lg
db = torch.empty(64 * n, dtype=torch.uint8, device="cuda")
ds = torch.empty(32 * n, dtype=torch.uint8, device="cuda")
ctx.gen_bases_dev(1, n, db.data_ptr())
ctx.gen_scalars_dev(2, n, ds.data_ptr())
hb = ctx.bases_from_device(db.data_ptr(), n)
if "--no-precompute" not in sys.argv:
    hb.precompute(-1)      # the bench.py configuration: per-Params window tables
for _ in range(2):
    r = ctx.msm_dev(hb, ds.data_ptr(), n)
w = h2a.fr_root_of_unity(lg)
for _ in range(2):
    ctx.ntt_dev(ds.data_ptr(), lg, w)
print("ok", bytes(r[:8]).hex())
hb.free()
ctx.close()
