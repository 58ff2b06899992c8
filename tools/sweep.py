#!/usr/bin/env python
"""Sweeps of BASELINE configs 1 and 2 on one GPU: G1 MSM 2^16..2^24 and Fr NTT k=16..24, inputs
resident in HBM, timed with CUDA events on the library stream.  One JSON line per size."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import halo2_aggregation_b200 as h2a


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--msm", default="16,18,20,22,24")
    ap.add_argument("--ntt", default="16,18,20,22,24")
    ap.add_argument("--windows", default="")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--algo", type=int, default=1)
    ap.add_argument("--precompute", default="", help="comma list of table window widths to time as well")
    args = ap.parse_args()
    ctx = h2a.Context(0)
    ctx.set_profiling(True)
    ctx.set_msm_algorithm(args.algo)
    stream = torch.cuda.ExternalStream(ctx.stream)

    def timeit(fn, reps):
        for _ in range(2):
            fn()
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        ctx.sync()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    msm_sizes = [int(x) for x in args.msm.split(",") if x]
    if msm_sizes:
        nmax = 1 << max(msm_sizes)
        db = torch.empty(64 * nmax, dtype=torch.uint8, device="cuda")
        ds = torch.empty(32 * nmax, dtype=torch.uint8, device="cuda")
        ctx.gen_bases_dev(1, nmax, db.data_ptr())
        ctx.gen_scalars_dev(2, nmax, ds.data_ptr())
        hb = ctx.bases_from_device(db.data_ptr(), nmax)
        for lg in msm_sizes:
            n = 1 << lg
            wins = [int(x) for x in args.windows.split(",") if x] or [0]
            for c in wins:
                ctx.set_msm_window(c)
                ms = timeit(lambda: ctx.msm_dev(hb, ds.data_ptr(), n), args.reps)
                print(json.dumps({"op": "msm_g1", "log_n": lg, "window": c, "ms": ms, "mpts_per_s": n / ms / 1e3,
                                  "phases_ms": dict(ctx.last_phases(0))}), flush=True)
            ctx.set_msm_window(0)
        for c in [int(x) for x in args.precompute.split(",") if x]:
            for lg in msm_sizes:
                n = 1 << lg
                hp = ctx.bases_from_device(db.data_ptr(), n)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); hp.precompute(c); e1.record(stream); ctx.sync(); torch.cuda.synchronize()
                ms = timeit(lambda: ctx.msm_dev(hp, ds.data_ptr(), n), args.reps)
                print(json.dumps({"op": "msm_g1_precomputed", "log_n": lg, "window": c, "ms": ms, "mpts_per_s": n / ms / 1e3,
                                  "precompute_ms": e0.elapsed_time(e1), "phases_ms": dict(ctx.last_phases(0))}), flush=True)
                hp.free()
        hb.free()
        del db, ds
    for k in [int(x) for x in args.ntt.split(",") if x]:
        n = 1 << k
        d = torch.empty(32 * n, dtype=torch.uint8, device="cuda")
        ctx.gen_scalars_dev(3, n, d.data_ptr())
        w = h2a.fr_root_of_unity(k)
        ms = timeit(lambda: ctx.ntt_dev(d.data_ptr(), k, w), args.reps)
        ph = dict(ctx.last_phases(1))
        kern_ms = sum(ph.values())
        print(json.dumps({"op": "ntt_fr", "log_n": k, "ms": ms, "kernel_ms": kern_ms, "melem_per_s": n / ms / 1e3,
                          "algo_gbs": 64 * n / (kern_ms * 1e-3) / 1e9, "phases_ms": ph}), flush=True)
        # the other transforms of EvaluationDomain: inverse (+ 1/n), coset forward / inverse (zeta-distribute convention)
        zeta = ctx.field_op(1, "to_mont", np.frombuffer((7).to_bytes(32, "little"), dtype=np.uint8)).copy()
        for name, inv, shift in (("intt_fr", True, None), ("coset_ntt_fr", False, zeta), ("coset_intt_fr", True, zeta)):
            ms = timeit(lambda: ctx.ntt_dev(d.data_ptr(), k, w, inverse=inv, coset_shift=shift), args.reps)
            kern_ms = sum(dict(ctx.last_phases(1)).values())
            print(json.dumps({"op": name, "log_n": k, "ms": ms, "kernel_ms": kern_ms, "melem_per_s": n / ms / 1e3}), flush=True)
        del d
        if k + 2 <= 24:   # coeff_to_extended n -> 4n through the host-buffer entry point (copies inside the time)
            coeffs = torch.empty(32 * n, dtype=torch.uint8).pin_memory()
            coeffs.numpy()[:] = 0
            coeffs.numpy()[::32] = 1
            ext = torch.empty(128 * n, dtype=torch.uint8).pin_memory()
            ms = timeit(lambda: ctx.coeff_to_extended(coeffs.numpy(), k, k + 2, zeta, out=ext.numpy()), max(2, args.reps // 2))
            print(json.dumps({"op": "coeff_to_extended_host", "log_n": k, "ext_log_n": k + 2, "ms": ms,
                              "note": "h2a_coeff_to_extended with host buffers: pinned host buffers, H2D of n and D2H of 4n elements inside the time"}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
