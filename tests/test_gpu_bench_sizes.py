"""GPU parity at the BENCHMARK's own configurations, compared directly with the CPU oracle (not with the CUDA path itself):
MSM 2^22 with the c=20 window tables bench.py times, the prover's configuration (2^20-point columns, c=17 tables, grouped
passes), NTT / coset NTT / coeff_to_extended at k=22, NTT at k=24, and the k=20 proof through the oracle's verifier.
Inputs are generated on the device (tests/test_gpu_parity.py::test_generators_match_oracle pins the generators to the
oracle's) and downloaded, so the oracle reads the very bytes the kernels read."""
import os
import sys

import numpy as np
import pytest

import halo2_aggregation_b200 as h2a
from oracle import plonk as pk
from oracle import pymodel as pm

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ctx():
    c = h2a.Context(0)
    yield c
    c.close()


def fr_bytes(v):
    return np.frombuffer(pm.fr_mont_bytes(v), dtype=np.uint8)


@pytest.mark.parametrize("algo", [1, 0], ids=["affine-tree", "xyzz-tasks"])
def test_msm_bench_config_matches_oracle(ctx, orc, algo):
    """bench.py's headline: 2^22 points, window tables c=20 — against the oracle's best_multiexp; also without tables,
    through host scalars (the e2e path: two point ranges on two lanes) and through the pipelined batch entry point."""
    n = 1 << 22
    ctx.set_msm_algorithm(algo)
    db, ds = ctx.dev_alloc(64 * n), ctx.dev_alloc(32 * n)
    ctx.gen_bases_dev(1, n, db)
    ctx.gen_scalars_dev(2, n, ds)
    bases, scalars = ctx.d2h(db, 64 * n), ctx.d2h(ds, 32 * n)
    want = bytes(orc.msm(bases, scalars))
    assert want != bytes(64)
    hb = ctx.bases_from_device(db, n)
    try:
        assert bytes(ctx.msm_dev(hb, ds, n)) == want                       # no tables (c = 16)
        hb.precompute(-1)                                                  # the benchmark's tables (c = 20)
        assert bytes(ctx.msm_dev(hb, ds, n)) == want
        assert bytes(ctx.msm(hb, scalars)) == want
        assert all(bytes(r) == want for r in ctx.msm_batch_dev(hb, [ds, ds, ds], [n, n, n]))
    finally:
        ctx.set_msm_algorithm(1)
        hb.free(); ctx.dev_free(db); ctx.dev_free(ds)


def test_msm_prover_config_matches_oracle(ctx, orc):
    """The prover's commitment configuration: columns of 2^20 scalars over tables of 17-bit windows, eight columns per
    pass — full-width columns, a 0/1 selector column and a small-value (range-checked) column, each against the oracle."""
    n = 1 << 20
    db = ctx.dev_alloc(64 * n)
    ctx.gen_bases_dev(11, n, db)
    bases = ctx.d2h(db, 64 * n)
    cols = []
    for j in range(6):
        d = ctx.dev_alloc(32 * n)
        ctx.gen_scalars_dev(20 + j, n, d)
        cols.append(d)
    rng = np.random.default_rng(5)
    small = np.zeros((n, 4), dtype=np.uint64); small[:, 0] = rng.integers(0, 1 << 16, n)
    sel = np.zeros((n, 4), dtype=np.uint64); sel[:, 0] = rng.integers(0, 2, n)
    for arr in (small, sel):
        d = ctx.dev_alloc(32 * n)
        ctx.h2d(d, ctx.field_op(1, "to_mont", arr.view(np.uint8).reshape(-1)))
        cols.append(d)
    want = [bytes(orc.msm(bases, ctx.d2h(d, 32 * n))) for d in cols]
    hb = ctx.bases_from_device(db, n)
    hb.precompute(17)
    try:
        ctx.set_msm_group(8, 2)
        got = ctx.msm_batch_dev(hb, cols, [n] * len(cols))
        assert [bytes(g) for g in got] == want
        ctx.set_msm_group(3, 2)
        got = ctx.msm_batch_dev(hb, cols, [n] * len(cols))
        assert [bytes(g) for g in got] == want
    finally:
        ctx.set_msm_group(8, 2)
        hb.free(); ctx.dev_free(db)
        for d in cols:
            ctx.dev_free(d)


@pytest.mark.parametrize("k", [22, 24])
def test_ntt_bench_sizes_match_oracle(ctx, orc, k):
    n = 1 << k
    w = h2a.fr_root_of_unity(k)
    d = ctx.dev_alloc(32 * n)
    ctx.gen_scalars_dev(300 + k, n, d)
    a = ctx.d2h(d, 32 * n)
    ctx.ntt_dev(d, k, w)
    got = ctx.d2h(d, 32 * n)
    assert bytes(got) == bytes(orc.fft(a, k, w))
    ctx.ntt_dev(d, k, w, inverse=True)
    assert bytes(ctx.d2h(d, 32 * n)) == bytes(a)
    if k == 22:     # coset transform and its inverse (the quotient's way back), host-pointer entry point too
        g = fr_bytes(7)
        ctx.ntt_dev(d, k, w, coset_shift=g)
        coset = ctx.d2h(d, 32 * n)
        assert bytes(coset) == bytes(orc.coeff_to_extended(a, k, k, g))
        assert bytes(ctx.ntt(a, k, w)) == bytes(got)
        ctx.extended_to_coeff_dev(d, k, g)
        assert bytes(ctx.d2h(d, 32 * n)) == bytes(a)
    ctx.dev_free(d)


def test_coeff_to_extended_bench_size_matches_oracle(ctx, orc):
    """The quotient path of a k=20 proof: 2^20 coefficients -> 2^22 coset evaluations, and back."""
    k, ext_k = 20, 22
    n, m = 1 << k, 1 << ext_k
    g = fr_bytes(7)
    d_in, d_out = ctx.dev_alloc(32 * n), ctx.dev_alloc(32 * m)
    ctx.gen_scalars_dev(91, n, d_in)
    coeffs = ctx.d2h(d_in, 32 * n)
    ctx.coeff_to_extended_dev(d_in, k, ext_k, g, d_out)
    ext = ctx.d2h(d_out, 32 * m)
    assert bytes(ext) == bytes(orc.coeff_to_extended(coeffs, k, ext_k, g))
    ctx.extended_to_coeff_dev(d_out, ext_k, g)
    back = ctx.d2h(d_out, 32 * m)
    assert bytes(back) == bytes(orc.extended_to_coeff(ext, ext_k, g))
    assert bytes(back[:32 * n]) == bytes(coeffs) and not back[32 * n:].any()
    ctx.dev_free(d_in); ctx.dev_free(d_out)


def test_k20_proof_accepted_by_oracle_verifier(ctx, orc):
    """The metric's first half: the k=20 aggregation-profile proof bench.py times is replayed by the oracle's verifier
    (restating VerifierChip::_verify_proof) and must satisfy the pairing relation; the library's verifier glue must
    return the oracle's (e, f, w, zw) bit for bit."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_extras
    import prove_bench

    class A:
        k, steps, lookups, precompute, world, rank = 20, 1, 9, prove_bench.PROVER_TABLE_BITS, 1, 0
    res = prove_bench.run(ctx, A)
    chk = res["_check"]
    ok, ores = bench_extras._oracle_accepts(dict(pk=pk, pm=pm), chk)
    assert ok and res["proof_verifies"]
    want = b"".join(pm.affine_bytes(ores[nm]) for nm in ("e", "f", "w", "zw"))
    assert bytes(res["_efwzw"]) == want
    # a flipped proof byte in an evaluation must be rejected
    bad = bytearray(chk["proof"]); bad[-200] ^= 1
    try:
        ok2, _ = bench_extras._oracle_accepts(dict(pk=pk, pm=pm), dict(chk, proof=bytes(bad)))
    except (AssertionError, ValueError):
        ok2 = False
    assert not ok2
