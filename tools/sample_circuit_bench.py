#!/usr/bin/env python
"""BASELINE config 0: the reference's sample circuit (`MyCircuit`, examples/simple-example.rs:316-391: one mul gate,
a u8 lookup, equality on instance / constant / advice) at its own size k = 9 (:561): KZG setup, key generation,
proof, and the off-circuit verify-accumulate, timed through the library.  The witness model is the one the parity
tests use (tests/circuits.py); proofs are checked with the pairing relation s*W == ZW + F + E.

  python tools/sample_circuit_bench.py [--k 9] [--proofs 64]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np

import halo2_aggregation_b200 as h2a
from prove_bench import R, Shape, canonical_bytes, fr, random_blinds

A, F, I = 0, 1, 2
OP_ADVICE, OP_FIXED, OP_NEG, OP_ADD, OP_MUL = 1, 2, 4, 5, 6


def my_circuit(ctx, k, a=200, b=13, constant=99):
    n, bf = 1 << k, 5
    s = Shape()
    s.k, s.bf, s.degree, s.num_instance, s.num_advice, s.num_fixed = k, bf, 5, 1, 2, 4
    s.advice_queries = [(0, 0), (1, 0), (0, 1)]
    s.fixed_queries = [(2, 0), (3, 0), (1, 0), (0, 0)]
    s.instance_queries = [(0, 0)]
    s.gates = [[(OP_FIXED, 2), (OP_ADVICE, 0), (OP_ADVICE, 1), (OP_MUL, 0), (OP_ADVICE, 2), (OP_NEG, 0), (OP_ADD, 0), (OP_MUL, 0)]]
    s.constants = []
    s.lookups = [([[(OP_FIXED, 0), (OP_ADVICE, 0), (OP_MUL, 0)]], [[(OP_FIXED, 1)]])]
    s.perm_columns = [(I, 0, 0), (F, 0, 3), (A, 0, 0), (A, 1, 1)]
    ab, absq = a * b, (a * b) ** 2
    c = constant * absq
    a0, a1 = np.zeros(n, dtype=object), np.zeros(n, dtype=object)
    fx = np.zeros((4, n), dtype=object)
    fx[3, :256] = np.arange(256)
    a0[0], fx[2, 0] = a, 1
    a0[1], fx[2, 1] = b, 1
    a0[2], fx[0, 0] = constant, constant
    fx[1, 3], a0[3], a1[3], a0[4] = 1, a, b, ab
    fx[1, 5], a0[5], a1[5], a0[6] = 1, ab, ab, absq
    fx[1, 7], a0[7], a1[7], a0[8] = 1, constant, absq, c
    rng = np.random.default_rng(1)
    for col in (a0, a1):
        for r in range(n - bf, n):
            col[r] = int(rng.integers(0, 1 << 62))
    inst = np.zeros(n, dtype=object)
    inst[0] = c
    to_b = lambda cols: fr(ctx, [int(v) for col in cols for v in col])
    # sigma from the copy cycles of synthesize(): (perm column position, row)
    cycles = [[(2, 0), (2, 3)], [(2, 1), (3, 3)], [(2, 2), (1, 0), (2, 7)], [(2, 4), (2, 5), (3, 5)], [(2, 6), (3, 7)], [(2, 8), (0, 0)]]
    delta = pow(7, 1 << 28, R)
    omega = pow(pow(7, (R - 1) >> 28, R), 1 << (28 - k), R)
    nxt = {}
    for cyc in cycles:
        for x, y in zip(cyc, cyc[1:] + cyc[:1]):
            nxt[x] = y
    om = [1]
    for _ in range(n - 1):
        om.append(om[-1] * omega % R)
    sig = [[pow(delta, nxt.get((j, r), (j, r))[0], R) * om[nxt.get((j, r), (j, r))[1]] % R for r in range(n)] for j in range(4)]
    return s, to_b([inst]), to_b([a0, a1]), to_b(fx), to_b(sig)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=9)
    ap.add_argument("--proofs", type=int, default=64)
    args = ap.parse_args()
    ctx = h2a.Context(0)
    seed = bytes([0x59, 0x62, 0xbe, 0x5d, 0x76, 0x3d, 0x31, 0x8d, 0x17, 0xdb, 0x37, 0x32, 0x54, 0x06, 0xbc, 0xe5])
    secret = h2a.xorshift_scalar(seed)                       # examples/simple-example.rs:584-589
    t0 = time.perf_counter(); g, gl = ctx.kzg_setup(args.k, secret); t_setup = time.perf_counter() - t0
    shape, inst_b, adv_b, fixed_b, sig_b = my_circuit(ctx, args.k)
    circ = h2a.Circuit(ctx, shape, np.zeros(0, np.uint8))
    t0 = time.perf_counter(); circ.set_keys(g, gl, fixed_b, sig_b, fr(ctx, [0xC0FFEE]), fr(ctx, [7])); t_keys = time.perf_counter() - t0
    circ.prove(inst_b, adv_b, random_blinds(ctx, circ.blinds_len(), 0))
    t0 = time.perf_counter()
    proofs, insts = [], []
    for p in range(args.proofs):
        proof, inst = circ.prove(inst_b, adv_b, random_blinds(ctx, circ.blinds_len(), p))
        proofs.append(proof); insts.append(inst)
    t_prove = time.perf_counter() - t0
    t0 = time.perf_counter(); res = circ.verify_batch(np.concatenate(insts), proofs); t_verify = time.perf_counter() - t0
    ok = True
    for i in (0, args.proofs - 1):
        e, f, w, zw = (res[i][64 * j:64 * j + 64] for j in range(4))
        ok &= bytes(ctx.msm_adhoc(w, secret)) == bytes(h2a.g1_sum(np.concatenate([zw, f, e])))
    print(json.dumps({"workload": "MyCircuit (examples/simple-example.rs:316-391) at k=%d, %d proofs" % (args.k, args.proofs),
                      "kzg_setup_s": t_setup, "keygen_s": t_keys, "prove_ms_per_proof": t_prove / args.proofs * 1e3,
                      "verify_accumulate_batch_ms": t_verify * 1e3, "proof_bytes": len(proofs[0]), "all_pairing_relations_hold": bool(ok)}))
    circ.free(); g.free(); gl.free(); ctx.close()


if __name__ == "__main__":
    main()
