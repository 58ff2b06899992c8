#!/usr/bin/env python
"""Row f4 measurement: witness cells of the non-native mul_var of a batch of aggregated proofs (37 per proof,
src/multiopen.rs:393-492 + src/vanishing.rs:181-187) in one launch.  Timing only: the cell-for-cell parity against the oracle lives in
tests/test_gpu_mulvar.py and in bench.py's `mulvar` block; here the products are cross-checked against the library's own MSM.
  python tools/mulvar_bench.py [--proofs 64] [--steps 3]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import halo2_aggregation_b200 as h2a


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--proofs", type=int, default=64)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    ctx = h2a.Context(0)
    m = 37 * args.proofs
    ln = ctx.mulvar_witness_len()
    d_aux = ctx.dev_alloc(64)
    ctx.gen_bases_dev(33, 1, d_aux)                     # any curve point serves as the auxiliary point
    aux = ctx.d2h(d_aux, 64)
    d_p, d_s, d_r, d_w = ctx.dev_alloc(64 * m), ctx.dev_alloc(32 * m), ctx.dev_alloc(64 * m), ctx.dev_alloc(32 * ln * m)
    ctx.gen_bases_dev(31, m, d_p)
    ctx.gen_scalars_dev(32, m, d_s)
    times = []
    for _ in range(args.steps + 1):
        ctx.sync()
        t0 = time.perf_counter()
        status = ctx.mulvar_witness_dev(d_p, d_s, m, aux, d_r, d_w)
        ctx.sync()
        times.append(time.perf_counter() - t0)
    best = min(times[1:])
    # the products against the library's own scalar multiplication (a one-term MSM) on a sample of entries
    pts, scal, res = ctx.d2h(d_p, 64 * m), ctx.d2h(d_s, 32 * m), ctx.d2h(d_r, 64 * m)
    ok = not status.any()
    for i in (0, m // 2, m - 1):
        ok = ok and bytes(ctx.msm_adhoc(pts[64 * i:64 * i + 64], scal[32 * i:32 * i + 32])) == bytes(res[64 * i:64 * i + 64])
    print(json.dumps({"metric": "mul_var witness generation", "proofs": args.proofs, "mul_var": m, "cells_per_mul_var": ln,
                      "value": best, "unit": "s", "higher_is_better": False, "all_s": times[1:], "mul_var_per_s": m / best,
                      "witness_bytes": 32 * ln * m, "write_gb_per_s": 32 * ln * m / best / 1e9,
                      "products_checked": "3 entries against h2a_msm_g1_adhoc" if ok else "FAILED"}))
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
