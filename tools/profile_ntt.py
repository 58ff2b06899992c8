#!/usr/bin/env python
"""Short fixed workload for ncu: 3 NTTs of 2^22 elements on device-resident data."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import halo2_aggregation_b200 as h2a

ctx = h2a.Context(0)
lg = 22
ds = torch.empty(32 << lg, dtype=torch.uint8, device="cuda")
ctx.gen_scalars_dev(2, 1 << lg, ds.data_ptr())
w = h2a.fr_root_of_unity(lg)
for _ in range(3):
    ctx.ntt_dev(ds.data_ptr(), lg, w)
print("ok")
ctx.close()
