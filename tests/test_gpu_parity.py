"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs (bit-exact), plus size-independent properties at the benchmark sizes."""
import random

import numpy as np
import pytest

import halo2_aggregation_b200 as h2a
from oracle import pymodel as pm

pytestmark = pytest.mark.gpu
P, R = pm.P, pm.R


@pytest.fixture(scope="module", params=[1, 0], ids=["affine-tree", "xyzz-tasks"])
def ctx(request):
    """Every test runs with both bucket-accumulation algorithms of the MSM (h2a_msm_set_algorithm)."""
    c = h2a.Context(0)
    c.set_msm_algorithm(request.param)
    yield c
    c.close()


def ints_to_bytes(vals):
    return np.frombuffer(b"".join(pm.le32(v) for v in vals), dtype=np.uint8)


def fr_bytes(v):
    return np.frombuffer(pm.fr_mont_bytes(v), dtype=np.uint8)


def edge(m):
    return [0, 1, 2, m - 1, m - 2, (m - 1) // 2, 1 << 253, (1 << 253) - 1, pm.MONT % m, (m - pm.MONT % m) % m,
            (1 << 32) - 1, 1 << 32, (1 << 64) - 1, (1 << 224) - 1]


# ------------------------------------------------------------------ field / curve arithmetic
@pytest.mark.parametrize("field,mod", [(0, P), (1, R)])
def test_field_ops(ctx, orc, field, mod):
    rng = random.Random(40 + field)
    a = edge(mod) + [rng.randrange(mod) for _ in range(5000)]
    b = list(reversed(edge(mod))) + [rng.randrange(mod) for _ in range(5000)]
    am, bm = orc.to_mont(field, ints_to_bytes(a)), orc.to_mont(field, ints_to_bytes(b))
    for op in ("add", "sub", "mul", "sqr", "neg"):
        assert bytes(ctx.field_op(field, op, am, bm)) == bytes(orc.field_op(field, op, am, bm)), op
    assert bytes(ctx.field_op(field, "inv", am[:32 * 64])) == bytes(orc.field_op(field, "inv", am[:32 * 64]))
    assert bytes(ctx.field_op(field, "inv_fast", am)) == bytes(orc.field_op(field, "inv", am))
    # values with long runs of zero / one bits, small values and their negatives (division-step corner cases)
    special = [1 << i for i in range(254)] + [(mod - (1 << i)) % mod for i in range(254)] + [(1 << i) - 1 for i in range(1, 254)]
    special = [v % mod for v in special] + list(range(0, 64)) + [mod - i for i in range(1, 64)]
    sm = orc.to_mont(field, ints_to_bytes(special))
    assert bytes(ctx.field_op(field, "inv_fast", sm)) == bytes(orc.field_op(field, "inv", sm))
    raw = ints_to_bytes(special)   # the same integers taken as Montgomery representatives
    assert bytes(ctx.field_op(field, "inv_fast", raw)) == bytes(orc.field_op(field, "inv", raw))
    # all pairs of edge values through mul
    e = edge(mod)
    aa = orc.to_mont(field, ints_to_bytes([x for x in e for _ in e]))
    bb = orc.to_mont(field, ints_to_bytes([y for _ in e for y in e]))
    assert bytes(ctx.field_op(field, "mul", aa, bb)) == bytes(orc.field_op(field, "mul", aa, bb))


def test_g1_ops(ctx, orc):
    n = 256
    a, b = orc.gen_bases(1, n), orc.gen_bases(2, n)
    a = a.copy(); b = b.copy()
    b[0:64] = a[0:64]                      # P + P
    neg = pm.affine_bytes(pm.g1_neg(pm.affine_from_bytes(a[64:128])))
    b[64:128] = np.frombuffer(neg, dtype=np.uint8)   # P + (-P)
    b[128:192] = 0                          # P + O
    a[192:256] = 0                          # O + P
    a[256:320] = 0; b[256:320] = 0          # O + O
    got = ctx.g1_op("add", a, b)
    for i in range(n):
        want = orc.g1_add(a[64 * i:64 * i + 64], b[64 * i:64 * i + 64])
        assert bytes(got[64 * i:64 * i + 64]) == bytes(want), i
    got = ctx.g1_op("dbl", a)
    for i in range(n):
        assert bytes(got[64 * i:64 * i + 64]) == bytes(orc.g1_add(a[64 * i:64 * i + 64], a[64 * i:64 * i + 64])), i
    got = ctx.g1_op("dbl_add", a, b)
    for i in range(n):
        d = orc.g1_add(a[64 * i:64 * i + 64], a[64 * i:64 * i + 64])
        assert bytes(got[64 * i:64 * i + 64]) == bytes(orc.g1_add(d, b[64 * i:64 * i + 64])), i


def test_generators_match_oracle(ctx, orc):
    n = 3000
    d = ctx.dev_alloc(64 * n)
    ctx.gen_bases_dev(11, n, d, first=5)
    assert bytes(ctx.d2h(d, 64 * n)) == bytes(orc.gen_bases(11, n, first=5))
    ctx.gen_scalars_dev(11, n, d, first=7)
    assert bytes(ctx.d2h(d, 32 * n)) == bytes(orc.gen_scalars(11, n, first=7))
    ctx.dev_free(d)


# ------------------------------------------------------------------ MSM
def test_msm_kat(ctx):
    pts = [pm.g1_mul(pm.G1, i) for i in range(1, 5)]
    bases = np.frombuffer(b"".join(pm.affine_bytes(p) for p in pts), dtype=np.uint8)
    scal = np.frombuffer(b"".join(pm.fr_mont_bytes(i) for i in range(1, 5)), dtype=np.uint8)
    want = (0x036083bfa420b15a4c11f66a3cffd55318b019feb45f833a876e93848625f5ae,
            0x2630c348c019c3edb74fe62a7e921361aae9621988223514d56ca8b36adc9e36)
    assert pm.affine_from_bytes(ctx.msm_adhoc(bases, scal)) == want
    b = ctx.upload_bases(bases)
    for c in (6, 9, 13, 16):
        ctx.set_msm_window(c)
        assert pm.affine_from_bytes(ctx.msm(b, scal)) == want
    ctx.set_msm_window(0)
    b.free()


@pytest.mark.parametrize("n", [1, 2, 31, 1000, 4097, 1 << 14])
def test_msm_random_matches_oracle(ctx, orc, n):
    bases, scal = orc.gen_bases(100 + n, n), orc.gen_scalars(200 + n, n)
    want = bytes(orc.msm(bases, scal))
    b = ctx.upload_bases(bases)
    assert bytes(ctx.msm(b, scal)) == want
    assert bytes(ctx.msm_adhoc(bases, scal)) == want
    b.free()


@pytest.mark.parametrize("c", [6, 7, 8, 10, 11, 12, 13, 15, 16, 17, 18, 20])
def test_msm_every_window_width(ctx, orc, c):
    n = 3000
    bases, scal = orc.gen_bases(7, n), orc.gen_scalars(8, n)
    want = bytes(orc.msm(bases, scal))
    b = ctx.upload_bases(bases)
    ctx.set_msm_window(c)
    try:
        assert bytes(ctx.msm(b, scal)) == want
    finally:
        ctx.set_msm_window(0)
        b.free()


def test_msm_adversarial_inputs(ctx, orc):
    n = 2048
    bases = orc.gen_bases(21, n).copy()
    b_pts = [bases[64 * i:64 * i + 64] for i in range(n)]
    cases = {
        "zeros": [0] * n,
        "ones": [1] * n,
        "r_minus_1": [R - 1] * n,
        "selector": [i & 1 for i in range(n)],
        "small16": [(i * 2654435761) & 0xffff for i in range(n)],
        "top_window": [R - 1 - i for i in range(n)],
        "half": [(R - 1) // 2 + (i % 3) for i in range(n)],
        "pow2": [1 << (i % 254) for i in range(n)],
        "window_edges": [((1 << 15) << (16 * (i % 15))) + (i % 2) for i in range(n)],
    }
    # one base repeated (forces P+P inside a bucket), P / -P pairs with equal scalars, identity bases
    rep = bases.copy()
    for i in range(0, 64):
        rep[64 * i:64 * i + 64] = b_pts[0]
    for i in range(64, 128, 2):
        neg = pm.affine_bytes(pm.g1_neg(pm.affine_from_bytes(b_pts[i])))
        rep[64 * (i + 1):64 * (i + 2)] = np.frombuffer(neg, dtype=np.uint8)
    rep[64 * 200:64 * 210] = 0
    rng = random.Random(5)
    mixed = [rng.randrange(R) for _ in range(n)]
    for i in range(0, 64):
        mixed[i] = mixed[0]
    for i in range(64, 128, 2):
        mixed[i + 1] = mixed[i]
    hb, hr = ctx.upload_bases(bases), ctx.upload_bases(rep)
    for name, sc in cases.items():
        sm = orc.to_mont(1, ints_to_bytes(sc))
        assert bytes(ctx.msm(hb, sm)) == bytes(orc.msm(bases, sm)), name
    sm = orc.to_mont(1, ints_to_bytes(mixed))
    assert bytes(ctx.msm(hr, sm)) == bytes(orc.msm(rep, sm))
    sm1 = orc.to_mont(1, ints_to_bytes([1] * n))
    assert bytes(ctx.msm(hr, sm1)) == bytes(orc.msm(rep, sm1))
    # offsets / empty
    assert bytes(ctx.msm(hb, sm[:32 * 100], offset=50)) == bytes(orc.msm(bases[64 * 50:64 * 150], sm[:32 * 100]))
    assert bytes(ctx.msm(hb, np.zeros(0, np.uint8))) == bytes(64)
    with pytest.raises(h2a.H2AError):
        ctx.msm(hb, sm, offset=1)
    hb.free(); hr.free()


def test_msm_skewed_scalars_large(ctx, orc):
    """Heavy buckets: every point in one bucket / two buckets / a 17-bit range (witness-like columns)."""
    n = 1 << 17
    bases = orc.gen_bases(55, n)
    hb = ctx.upload_bases(bases)
    rng = random.Random(3)
    cases = {
        "all_ones": np.frombuffer(pm.fr_mont_bytes(1) * n, dtype=np.uint8),
        "all_same_big": np.frombuffer(pm.fr_mont_bytes(R - 12345) * n, dtype=np.uint8),
        "bits": orc.to_mont(1, ints_to_bytes([rng.getrandbits(1) for _ in range(n)])),
        "limbs17": orc.to_mont(1, ints_to_bytes([rng.getrandbits(17) for _ in range(n)])),
        "mostly_zero": orc.to_mont(1, ints_to_bytes([rng.randrange(R) if i % 64 == 0 else 0 for i in range(n)])),
    }
    for name, sm in cases.items():
        assert bytes(ctx.msm(hb, sm)) == bytes(orc.msm(bases, sm)), name
    for c in (8, 12, 14):     # window widths whose top window holds only a few bits
        ctx.set_msm_window(c)
        sm = orc.gen_scalars(56, n)
        assert bytes(ctx.msm(hb, sm)) == bytes(orc.msm(bases, sm)), c
    ctx.set_msm_window(0)
    hb.free()


@pytest.mark.parametrize("c", [11, 13, 16, 20])
def test_msm_precomputed_tables(ctx, orc, c):
    """h2a_bases_precompute: same group element, bit for bit, for every kind of input."""
    n = 6000
    bases = orc.gen_bases(61, n).copy()
    bases[64 * 17:64 * 18] = 0                                   # an identity base
    bases[64 * 30:64 * 31] = bases[64 * 29:64 * 30]              # a repeated base
    hb = ctx.upload_bases(bases).precompute(c)
    rng = random.Random(c)
    cases = {
        "uniform": orc.gen_scalars(62, n),
        "ones": np.frombuffer(pm.fr_mont_bytes(1) * n, dtype=np.uint8),
        "r_minus_1": np.frombuffer(pm.fr_mont_bytes(R - 1) * n, dtype=np.uint8),
        "bits": orc.to_mont(1, ints_to_bytes([rng.getrandbits(1) for _ in range(n)])),
        "zeros": np.zeros(32 * n, np.uint8),
    }
    for name, sm in cases.items():
        assert bytes(ctx.msm(hb, sm)) == bytes(orc.msm(bases, sm)), name
    sm = cases["uniform"]
    assert bytes(ctx.msm(hb, sm[:32 * 1000], offset=2500)) == bytes(orc.msm(bases[64 * 2500:64 * 3500], sm[:32 * 1000]))
    # an explicit different window falls back to the plain path; dropping the tables too
    ctx.set_msm_window(9)
    assert bytes(ctx.msm(hb, sm)) == bytes(orc.msm(bases, sm))
    ctx.set_msm_window(0)
    hb.precompute(0)
    assert bytes(ctx.msm(hb, sm)) == bytes(orc.msm(bases, sm))
    with pytest.raises(h2a.H2AError):
        hb.precompute(25)
    hb.free()


def test_msm_batch_and_dev(ctx, orc):
    n = 5000
    bases = orc.gen_bases(31, n)
    cols = [orc.gen_scalars(40 + j, n - 100 * j) for j in range(3)]
    hb = ctx.upload_bases(bases)
    got = ctx.msm_batch(hb, cols)
    for j, col in enumerate(cols):
        assert bytes(got[j]) == bytes(orc.msm(bases[:64 * (n - 100 * j)], col))
    d = ctx.dev_alloc(32 * n)
    ctx.h2d(d, cols[0])
    assert bytes(ctx.msm_dev(hb, d, n)) == bytes(got[0])
    ctx.dev_free(d); hb.free()


def test_msm_pipelined_batches(ctx, orc):
    """h2a_msm_g1_batch / _batch_dev alternate columns over two lanes: same results as one call per column."""
    n = 1 << 15
    bases = orc.gen_bases(71, n)
    hb = ctx.upload_bases(bases)
    cols = [orc.gen_scalars(80 + j, n - 37 * j) for j in range(5)] + [np.zeros(0, np.uint8)]
    want = [bytes(orc.msm(bases[:64 * (c.size // 32)], c)) if c.size else bytes(64) for c in cols]
    got = ctx.msm_batch(hb, cols)
    assert [bytes(g) for g in got] == want
    hb.precompute(16)
    got = ctx.msm_batch(hb, cols)
    assert [bytes(g) for g in got] == want
    dptrs = []
    for c in cols[:5]:
        d = ctx.dev_alloc(max(c.size, 32)); ctx.h2d(d, c); dptrs.append(d)
    got = ctx.msm_batch_dev(hb, dptrs, [c.size // 32 for c in cols[:5]])
    assert [bytes(g) for g in got] == want[:5]
    # a single call after a batch still works (no pending state left behind)
    assert bytes(ctx.msm(hb, cols[0])) == want[0]
    for d in dptrs:
        ctx.dev_free(d)
    hb.free()


@pytest.mark.parametrize("c,group", [(16, 8), (12, 3), (16, 64)])
def test_msm_grouped_columns(ctx, orc, c, group):
    """Columns of equal length over bases with tables are committed several per pass (one bucket set per column):
    every column still gets its own best_multiexp value, whatever its scalars look like and however the batch is cut."""
    n = 1 << 13
    r_minus_1 = (0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001 - 1).to_bytes(32, "little")
    raw = lambda b: orc.to_mont(1, np.frombuffer(b * n, dtype=np.uint8))
    small = np.zeros((n, 32), np.uint8)
    small[:, :2] = np.random.default_rng(5).integers(0, 256, size=(n, 2), dtype=np.uint8)
    sel = np.zeros((n, 32), np.uint8)
    sel[::3, 0] = 1
    cols = [orc.gen_scalars(300, n), np.zeros(32 * n, np.uint8), raw((1).to_bytes(32, "little")), raw(r_minus_1),
            orc.to_mont(1, small.reshape(-1)), orc.to_mont(1, sel.reshape(-1)), orc.gen_scalars(301, n), orc.gen_scalars(302, n),
            orc.gen_scalars(303, n)]
    bases = orc.gen_bases(310, n)
    bases[64 * 5:64 * 6] = 0                              # an identity among the bases
    bases[64 * 8:64 * 9] = bases[64 * 7:64 * 8]           # a repeated base (P + P inside a bucket)
    want = [bytes(orc.msm(bases, col)) for col in cols]
    hb = ctx.upload_bases(bases).precompute(c)
    dptrs = []
    for col in cols:
        d = ctx.dev_alloc(col.size); ctx.h2d(d, col); dptrs.append(d)
    ctx.set_msm_group(group, 2)
    try:
        for m in (len(cols), 4, 2):
            got = ctx.msm_batch_dev(hb, dptrs[:m], [n] * m)
            assert [bytes(g) for g in got] == want[:m]
        # a single call afterwards (no pending state left behind), and the ungrouped path
        assert bytes(ctx.msm_dev(hb, dptrs[0], n)) == want[0]
        ctx.set_msm_group(1, 1)
        got = ctx.msm_batch_dev(hb, dptrs, [n] * len(cols))
        assert [bytes(g) for g in got] == want
    finally:
        ctx.set_msm_group(8, 2)
        for d in dptrs:
            ctx.dev_free(d)
        hb.free()


def test_msm_deep_tree_two_streams(ctx, orc):
    """2^20 points over tables of width 16: 512 entries per bucket, five rounds of the batched-affine tree.  The two halves
    of the tree run on two unordered streams; with a point array shared between them a half that got a round ahead
    overwrote points the other was still reading (wrong sums now and then).  Each half owns its arrays now: every repeat
    must give the oracle's value."""
    n = 1 << 20
    bases = orc.gen_bases(910, n)
    cols = [orc.gen_scalars(920 + j, n) for j in range(3)]
    want = [bytes(orc.msm(bases, c)) for c in cols]
    hb = ctx.upload_bases(bases).precompute(16)
    try:
        for rep in range(4):
            assert [bytes(ctx.msm(hb, c)) for c in cols] == want, rep
    finally:
        hb.free()


def test_msm_linearity_at_bench_size(ctx):
    """2^22 points (BASELINE metric size): MSM(s, B) over [0,n) equals the sum of MSMs over two halves,
    and MSM with all-one scalars over the first 2^16 bases equals the plain point sum."""
    n = 1 << 22
    db, ds = ctx.dev_alloc(64 * n), ctx.dev_alloc(32 * n)
    ctx.gen_bases_dev(1, n, db)
    ctx.gen_scalars_dev(2, n, ds)
    hb = ctx.bases_from_device(db, n)
    full = ctx.msm_dev(hb, ds, n)
    half = n // 2
    lo = ctx.msm_dev(hb, ds, half)
    hi = ctx.msm_dev(hb, ds + 32 * half, half, offset=half)
    assert bytes(h2a.g1_sum(np.concatenate([lo, hi]))) == bytes(full)
    assert bytes(full) != bytes(64)
    # a different window width gives the same group element
    ctx.set_msm_window(14)
    assert bytes(ctx.msm_dev(hb, ds, n)) == bytes(full)
    ctx.set_msm_window(0)
    # precomputed window tables (the benchmark configuration) give the same element
    hb.precompute(20)
    assert bytes(ctx.msm_dev(hb, ds, n)) == bytes(full)
    assert bytes(ctx.msm_dev(hb, ds + 32 * half, half, offset=half)) == bytes(hi)
    # host scalars: the call cuts the MSM into point ranges whose copies overlap compute (h2a_msm_set_host_split)
    host_scalars = ctx.d2h(ds, 32 * n)
    for pieces in (2, 3, 1):
        ctx.set_msm_host_split(pieces)
        assert bytes(ctx.msm(hb, host_scalars)) == bytes(full)
        assert bytes(ctx.msm(hb, host_scalars[32 * half:], offset=half)) == bytes(hi)
    ctx.set_msm_host_split(2)
    m = 1 << 16
    ones = np.frombuffer(pm.fr_mont_bytes(1) * m, dtype=np.uint8)
    pts = ctx.d2h(db, 64 * m)
    assert bytes(ctx.msm(hb, ones)) == bytes(h2a.g1_sum(pts))
    hb.free(); ctx.dev_free(db); ctx.dev_free(ds)


def test_msm_largest_baseline_size(ctx):
    """2^24 points (the top of BASELINE's sweep): the MSM equals the sum of the MSMs over its four quarters, with and
    without window tables, and a host-scalar call (split into point ranges) gives the same element."""
    n = 1 << 24
    db, ds = ctx.dev_alloc(64 * n), ctx.dev_alloc(32 * n)
    ctx.gen_bases_dev(5, n, db)
    ctx.gen_scalars_dev(6, n, ds)
    hb = ctx.bases_from_device(db, n)
    q = n // 4
    parts = [ctx.msm_dev(hb, ds + 32 * q * i, q, offset=q * i) for i in range(4)]
    want = bytes(h2a.g1_sum(np.concatenate(parts)))
    assert want != bytes(64)
    assert bytes(ctx.msm_dev(hb, ds, n)) == want
    hb.precompute(20)
    assert bytes(ctx.msm_dev(hb, ds, n)) == want
    host_scalars = ctx.d2h(ds, 32 * n)
    assert bytes(ctx.msm(hb, host_scalars)) == want
    hb.free(); ctx.dev_free(db); ctx.dev_free(ds)


# ------------------------------------------------------------------ NTT
@pytest.mark.parametrize("k", [1, 2, 3, 5, 8, 10, 11, 12, 13, 16, 20, 21])
def test_ntt_matches_oracle(ctx, orc, k):
    a = orc.gen_scalars(300 + k, 1 << k)
    w = orc.fr_root_of_unity(k)
    want = orc.fft(a, k, w)
    got = ctx.ntt(a, k, w)
    assert bytes(got) == bytes(want)
    winv = orc.field_op(1, "inv", w)
    assert bytes(ctx.ntt(got, k, w, inverse=True)) == bytes(a)
    assert bytes(ctx.ntt(a, k, w, inverse=True)) == bytes(orc.ifft(a, k, winv))


def test_ntt_small_definition(ctx):
    rng = random.Random(1)
    for k in (1, 2, 4, 6):
        a = [rng.randrange(R) for _ in range(1 << k)]
        w = pm.omega_for(k)
        got = ctx.ntt(np.frombuffer(b"".join(pm.fr_mont_bytes(v) for v in a), dtype=np.uint8), k, fr_bytes(w))
        assert [pm.fr_from_mont_bytes(got[32 * i:32 * i + 32]) for i in range(1 << k)] == pm.ntt(a, w)


@pytest.mark.parametrize("k,ext_k", [(4, 6), (9, 11), (10, 12), (12, 14), (16, 18)])
def test_coset_extension_matches_oracle(ctx, orc, k, ext_k):
    coeffs = orc.gen_scalars(500 + k, 1 << k)
    g = fr_bytes(7)
    want = orc.coeff_to_extended(coeffs, k, ext_k, g)
    got = ctx.coeff_to_extended(coeffs, k, ext_k, g)
    assert bytes(got) == bytes(want)
    back = ctx.extended_to_coeff(got, ext_k, g)
    assert bytes(back) == bytes(orc.extended_to_coeff(want, ext_k, g))
    assert bytes(back[:32 << k]) == bytes(coeffs) and not back[32 << k:].any()


def test_coset_extension_on_device_pointers(ctx, orc):
    """h2a_coeff_to_extended_dev / h2a_extended_to_coeff_dev: same bytes as the host-buffer entry points, round trip."""
    k, ext_k = 11, 13
    n, m = 1 << k, 1 << ext_k
    coeffs = orc.gen_scalars(77, n)
    zeta = fr_bytes(7)
    want = ctx.coeff_to_extended(coeffs, k, ext_k, zeta)
    d_in, d_out = ctx.dev_alloc(32 * n), ctx.dev_alloc(32 * m)
    ctx.h2d(d_in, coeffs)
    ctx.coeff_to_extended_dev(d_in, k, ext_k, zeta, d_out)
    assert bytes(ctx.d2h(d_out, 32 * m)) == bytes(want)
    assert bytes(ctx.d2h(d_in, 32 * n)) == bytes(coeffs)          # the input is left alone
    ctx.extended_to_coeff_dev(d_out, ext_k, zeta)
    back = ctx.d2h(d_out, 32 * m)
    assert bytes(back[:32 * n]) == bytes(coeffs) and bytes(back[32 * n:]) == bytes(32 * (m - n))
    with pytest.raises(h2a.H2AError):
        ctx.coeff_to_extended_dev(d_out, k, ext_k, zeta, d_out)       # overlapping buffers are refused
    ctx.dev_free(d_in); ctx.dev_free(d_out)


def test_ntt_roundtrip_and_linearity_at_bench_size(ctx, orc):
    k = 22
    n = 1 << k
    w = h2a.fr_root_of_unity(k)
    d = ctx.dev_alloc(32 * n)
    ctx.gen_scalars_dev(77, n, d)
    a = ctx.d2h(d, 32 * n)
    ctx.ntt_dev(d, k, w)
    fa = ctx.d2h(d, 32 * n)
    # spot-check outputs against the definition via Horner on the oracle side: A[i] = poly(omega^i)
    # (checked through the inverse instead: iNTT(NTT(a)) == a, and NTT(a)[0] == sum a)
    ctx.ntt_dev(d, k, w, inverse=True)
    assert bytes(ctx.d2h(d, 32 * n)) == bytes(a)
    s = np.zeros(32, np.uint8)
    chunks = a.reshape(-1, 32)
    acc = 0
    # sum of all inputs mod r (Montgomery form is linear)
    vals = chunks.view(np.uint64).reshape(-1, 4)
    tot = [int(vals[:, j].astype(object).sum()) for j in range(4)]
    acc = sum(t << (64 * j) for j, t in enumerate(tot)) % R
    assert pm.from_le(fa[:32]) == acc
    ctx.dev_free(d)


def test_evaluation_domain_mirror(ctx, orc):
    dom = h2a.EvaluationDomain(ctx, j=5, k=8, coset_shift=fr_bytes(7))
    assert dom.extended_k == 10 and dom.get_quotient_poly_degree() == 4
    a = orc.gen_scalars(1, 1 << 8)
    assert bytes(dom.coeff_to_lagrange(dom.lagrange_to_coeff(a))) == bytes(a)
    ext = dom.coeff_to_extended(a)
    assert bytes(dom.extended_to_coeff(ext)[:32 << 8]) == bytes(a)


# ------------------------------------------------------------------ verifier glue
def make_proof(orc, seed, rots):
    nq = len(rots)
    commitments = orc.gen_bases(1000 + seed, nq)
    evals = orc.gen_scalars(2000 + seed, nq)
    ws = orc.gen_bases(3000 + seed, len(set(rots)))
    xuv = orc.gen_scalars(4000 + seed, 3)
    return dict(commitments=commitments, rotations=rots, evals=evals, ws=ws, x=xuv[:32], u=xuv[32:64], v=xuv[64:96])


ROTS = [0, 0, -1, 1, 0, -6, 1, 0, 0, -1, 0, 0, 0, 1, -1, 0, 0, 0, 0, 0, 0, 0, 0, 0]


def test_verify_accumulate_matches_oracle(ctx, orc):
    omega = orc.fr_root_of_unity(9)
    g1 = np.frombuffer(pm.affine_bytes(pm.G1), dtype=np.uint8)
    p = make_proof(orc, 1, ROTS)
    want = orc.gwc_accumulate(p["commitments"], p["rotations"], p["evals"], p["ws"], p["x"], p["u"], p["v"], omega, g1)
    got = ctx.verify_accumulate(p["commitments"], p["rotations"], p["evals"], p["ws"], p["x"], p["u"], p["v"], omega, g1)
    assert bytes(got) == bytes(want)
    # single rotation set, single query
    p1 = make_proof(orc, 2, [0])
    want = orc.gwc_accumulate(p1["commitments"], p1["rotations"], p1["evals"], p1["ws"], p1["x"], p1["u"], p1["v"], omega, g1)
    got = ctx.verify_accumulate(p1["commitments"], p1["rotations"], p1["evals"], p1["ws"], p1["x"], p1["u"], p1["v"], omega, g1)
    assert bytes(got) == bytes(want)
    with pytest.raises(h2a.H2AError):
        ctx.verify_accumulate(p["commitments"], p["rotations"], p["evals"], p["ws"][:64], p["x"], p["u"], p["v"], omega, g1)


def test_verify_accumulate_batch_of_64(ctx, orc):
    omega = orc.fr_root_of_unity(9)
    g1 = np.frombuffer(pm.affine_bytes(pm.G1), dtype=np.uint8)
    proofs = [make_proof(orc, s, ROTS) for s in range(64)]
    got = ctx.verify_accumulate_batch(proofs, omega, g1)
    for s in (0, 1, 17, 63):
        p = proofs[s]
        want = orc.gwc_accumulate(p["commitments"], p["rotations"], p["evals"], p["ws"], p["x"], p["u"], p["v"], omega, g1)
        assert bytes(got[s]) == bytes(want)


def test_fold_h_matches_oracle(ctx, orc):
    hs = orc.gen_bases(9, 4)
    xn = orc.gen_scalars(9, 1)
    assert bytes(ctx.fold_h(hs, xn)) == bytes(orc.fold_h(hs, xn))
    assert bytes(ctx.fold_h(hs[:64], xn)) == bytes(hs[:64])
