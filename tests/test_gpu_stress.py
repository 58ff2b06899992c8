"""Cross-stream / cross-context stress (what compute-sanitizer's racecheck would look for, which is closed on this pool —
profiles/r2_compute_sanitizer_closed.log): two contexts on one GPU work at the same time from two host threads — one commits
grouped batches of columns (two lanes, two streams each, the batched-affine tree), the other writes proofs (transform lane +
commitment lanes) — and every result must stay bit-identical with the oracle's, call after call.  A hazard like the
two-stream race of round 1 (a tree half overwriting points the other half still read) shows up here as a wrong commitment."""
import threading

import numpy as np
import pytest

import circuits
import halo2_aggregation_b200 as h2a
from oracle import plonk as pk
from oracle import pymodel as pm

pytestmark = pytest.mark.gpu


def frs_bytes(vals):
    return np.frombuffer(b"".join(pm.fr_mont_bytes(v) for v in vals), dtype=np.uint8) if vals else np.zeros(0, np.uint8)


def cols_bytes(cols):
    return np.frombuffer(b"".join(pm.fr_mont_bytes(v) for col in cols for v in col), dtype=np.uint8)


def test_two_contexts_commit_and_prove_concurrently(orc):
    ctx_a, ctx_b = h2a.Context(0), h2a.Context(0)
    errors = []
    # --- context A: grouped commitments over tables (deep tree at 2^18 with 16-bit tables) and plain batches
    n = 1 << 18
    bases = orc.gen_bases(301, n)
    cols = [orc.gen_scalars(310 + j, n) for j in range(6)]
    want_cols = [bytes(orc.msm(bases, c)) for c in cols]
    hb = ctx_a.upload_bases(bases)
    hb.precompute(16)
    dptrs = []
    for c in cols:
        d = ctx_a.dev_alloc(c.size); ctx_a.h2d(d, c); dptrs.append(d)
    # --- context B: proofs of the sample circuit at the reference's size
    c = circuits.my_circuit(k=9, table_bits=7)
    params, keys = circuits.setup(orc, c)
    want_proof, _ = pk.create_proof(orc, params, c["shape"], keys, c["instance"], c["advice"], seed=5)
    g, gl = ctx_b.upload_bases(params.g), ctx_b.upload_bases(params.g_lagrange)
    circ = h2a.Circuit(ctx_b, c["shape"], frs_bytes(c["shape"].constants))
    circ.set_keys(g, gl, cols_bytes(c["fixed"]), cols_bytes(keys.sigmas), frs_bytes([keys.vk_hash]), frs_bytes([c["shape"].coset_shift]))
    inst_b, adv_b, blinds = cols_bytes(c["instance"]), cols_bytes(c["advice"]), frs_bytes(pk.blinds_buffer(c["shape"], 5))

    def committer():
        try:
            for it in range(12):
                got = ctx_a.msm_batch_dev(hb, dptrs, [n] * len(dptrs))
                if [bytes(x) for x in got] != want_cols:
                    errors.append("commitments differ in iteration %d" % it)
                    return
                if bytes(ctx_a.msm(hb, cols[it % 6])) != want_cols[it % 6]:
                    errors.append("single MSM differs in iteration %d" % it)
                    return
        except Exception as e:   # noqa: BLE001
            errors.append("committer: %r" % (e,))

    def prover():
        try:
            for it in range(25):
                proof, _ = circ.prove(inst_b, adv_b, blinds)
                if proof != want_proof:
                    errors.append("proof differs in iteration %d" % it)
                    return
        except Exception as e:   # noqa: BLE001
            errors.append("prover: %r" % (e,))

    ts = [threading.Thread(target=committer), threading.Thread(target=prover)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors
    for d in dptrs:
        ctx_a.dev_free(d)
    hb.free(); circ.free(); g.free(); gl.free()
    ctx_a.close(); ctx_b.close()
