"""Pins the C++ oracle (oracle/oracle.cpp) to the Python big-int model, the known answers of
SURVEY.md App. A, and Python's hashlib.blake2b.  CPU only."""
import hashlib
import random

import numpy as np
import pytest

from oracle import pymodel as pm

P, R = pm.P, pm.R


def ints_to_bytes(vals):
    return np.frombuffer(b"".join(pm.le32(v) for v in vals), dtype=np.uint8)


def bytes_to_ints(b):
    b = bytes(b)
    return [pm.from_le(b[i:i + 32]) for i in range(0, len(b), 32)]


EDGE = lambda m: [0, 1, 2, m - 1, m - 2, (m - 1) // 2, (1 << 253), (1 << 253) - 1, pm.MONT % m, (m - pm.MONT % m) % m]


@pytest.mark.parametrize("field,mod", [(0, P), (1, R)])
def test_field_ops_match_bigint(orc, field, mod):
    rng = random.Random(1234 + field)
    a = EDGE(mod) + [rng.randrange(mod) for _ in range(200)]
    b = list(reversed(EDGE(mod))) + [rng.randrange(mod) for _ in range(200)]
    am = orc.to_mont(field, ints_to_bytes(a))
    bm = orc.to_mont(field, ints_to_bytes(b))
    assert bytes_to_ints(am) == [pm.to_mont(v, mod) for v in a]
    assert bytes_to_ints(orc.from_mont(field, am)) == a
    exp = {
        "add": [(x + y) % mod for x, y in zip(a, b)],
        "sub": [(x - y) % mod for x, y in zip(a, b)],
        "mul": [(x * y) % mod for x, y in zip(a, b)],
        "sqr": [(x * x) % mod for x in a],
        "neg": [(-x) % mod for x in a],
        "inv": [pow(x, -1, mod) if x else 0 for x in a],
    }
    for op, want in exp.items():
        got = bytes_to_ints(orc.from_mont(field, orc.field_op(field, op, am, bm)))
        assert got == want, op


def test_roots_of_unity(orc):
    for k in (1, 4, 9, 20, 23, 28):
        w = bytes_to_ints(orc.from_mont(1, orc.fr_root_of_unity(k)))[0]
        assert w == pm.omega_for(k)
        assert pow(w, 1 << k, R) == 1 and pow(w, 1 << (k - 1), R) == R - 1
    assert pm.FR_ROOT == 0x03ddb9f5166d18b798865ea93dd31f743215cf6dd39329c8d34f1ed960c37c9c


KAT = {
    2: (0x030644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd3, 0x15ed738c0e0a7c92e7845f96b2ae9c0a68a6a449e3538fc7ff3ebf7a5a18a2c4),
    3: (0x0769bf9ac56bea3ff40232bcb1b6bd159315d84715b8e679f2d355961915abf0, 0x2ab799bee0489429554fdb7c8d086475319e63b40b9c5b57cdf1ff3dd9fe2261),
    7: (0x17072b2ed3bb8d759a5325f477629386cb6fc6ecb801bd76983a6b86abffe078, 0x168ada6cd130dd52017bb54bfa19377aadfe3bf05d18f41b77809f7f60d4af9e),
    30: (0x036083bfa420b15a4c11f66a3cffd55318b019feb45f833a876e93848625f5ae, 0x2630c348c019c3edb74fe62a7e921361aae9621988223514d56ca8b36adc9e36),
}


def gen_affine():
    return np.frombuffer(pm.affine_bytes(pm.G1), dtype=np.uint8)


def fr_bytes(v):
    return np.frombuffer(pm.fr_mont_bytes(v), dtype=np.uint8)


def test_g1_known_answers(orc):
    g = gen_affine()
    for k, want in KAT.items():
        assert pm.g1_mul(pm.G1, k) == want
        assert pm.affine_from_bytes(orc.g1_mul(g, fr_bytes(k))) == want
    assert pm.affine_from_bytes(orc.g1_mul(g, fr_bytes(R - 1))) == (1, P - 2)
    assert pm.affine_from_bytes(orc.g1_mul(g, fr_bytes(0))) is None
    # P + P, P + (-P), P + O
    two = orc.g1_add(g, g)
    assert pm.affine_from_bytes(two) == KAT[2]
    neg = np.frombuffer(pm.affine_bytes(pm.g1_neg(pm.G1)), dtype=np.uint8)
    assert pm.affine_from_bytes(orc.g1_add(g, neg)) is None
    assert pm.affine_from_bytes(orc.g1_add(g, np.zeros(64, np.uint8))) == pm.G1


def test_msm_kat_and_bigint(orc):
    # sum_{i=1..4} i*(iG) = 30G
    pts = [pm.g1_mul(pm.G1, i) for i in range(1, 5)]
    bases = np.frombuffer(b"".join(pm.affine_bytes(p) for p in pts), dtype=np.uint8)
    scal = np.frombuffer(b"".join(pm.fr_mont_bytes(i) for i in range(1, 5)), dtype=np.uint8)
    for naive in (False, True):
        for th in (1, 3):
            assert pm.affine_from_bytes(orc.msm(bases, scal, threads=th, naive=naive)) == KAT[30]
    rng = random.Random(7)
    n = 70
    pts = [pm.g1_mul(pm.G1, rng.randrange(1, R)) for _ in range(n)]
    sc = [rng.randrange(R) for _ in range(n)]
    sc[3] = 0; sc[5] = R - 1; sc[6] = 1; pts[8] = pts[7]; pts[10] = pm.g1_neg(pts[9]); sc[10] = sc[9]
    bases = np.frombuffer(b"".join(pm.affine_bytes(p) for p in pts), dtype=np.uint8)
    scal = np.frombuffer(b"".join(pm.fr_mont_bytes(s) for s in sc), dtype=np.uint8)
    want = pm.msm(sc, pts)
    for th in (1, 4, 8):
        assert pm.affine_from_bytes(orc.msm(bases, scal, threads=th)) == want
    assert pm.affine_from_bytes(orc.msm(bases, scal, naive=True)) == want


def test_gen_inputs(orc):
    b = orc.gen_bases(42, 300)
    assert orc.g1_on_curve(b)
    assert bytes(orc.gen_bases(42, 100, first=200)) == bytes(b[64 * 200:])
    assert bytes(orc.gen_bases(42, 300, threads=1)) == bytes(b)
    pts = [pm.affine_from_bytes(b[64 * i:64 * i + 64]) for i in range(300)]
    assert all(pm.on_curve(p) and p is not None for p in pts)
    assert len(set(pts)) == 300
    s = orc.gen_scalars(42, 1000)
    assert all(v < R for v in bytes_to_ints(s))
    assert bytes(orc.gen_scalars(42, 500, first=500)) == bytes(s[32 * 500:])
    # both y signs appear
    assert len({p[1] & 1 for p in pts}) == 2


def test_msm_pippenger_vs_naive_generated(orc):
    n = 3000
    bases, scal = orc.gen_bases(5, n), orc.gen_scalars(5, n)
    assert bytes(orc.msm(bases, scal)) == bytes(orc.msm(bases, scal, naive=True))
    assert bytes(orc.msm(bases, scal, threads=1)) == bytes(orc.msm(bases, scal, threads=5))


@pytest.mark.parametrize("k", [1, 2, 5, 8])
def test_fft_matches_definition(orc, k):
    rng = random.Random(k)
    n = 1 << k
    a = [rng.randrange(R) for _ in range(n)]
    am = np.frombuffer(b"".join(pm.fr_mont_bytes(v) for v in a), dtype=np.uint8)
    w = pm.omega_for(k)
    want = pm.ntt(a, w)
    for th in (1, 2, 8):
        got = bytes_to_ints(orc.from_mont(1, orc.fft(am, k, fr_bytes(w), threads=th)))
        assert got == want
    back = orc.ifft(orc.fft(am, k, fr_bytes(w)), k, fr_bytes(pow(w, -1, R)))
    assert bytes(back) == bytes(am)


def test_fft_parallel_equals_serial_large(orc):
    k = 14
    a = orc.gen_scalars(9, 1 << k)
    w = orc.fr_root_of_unity(k)
    assert bytes(orc.fft(a, k, w, threads=1)) == bytes(orc.fft(a, k, w, threads=8))


def test_coset_extension_roundtrip(orc):
    k, ext_k = 6, 8
    rng = random.Random(3)
    coeffs = [rng.randrange(R) for _ in range(1 << k)]
    cm = np.frombuffer(b"".join(pm.fr_mont_bytes(v) for v in coeffs), dtype=np.uint8)
    g = 7
    ext = orc.coeff_to_extended(cm, k, ext_k, fr_bytes(g))
    got = bytes_to_ints(orc.from_mont(1, ext))
    w = pm.omega_for(ext_k)
    for i in (0, 1, 17, 255):
        assert got[i] == pm.poly_eval(coeffs, g * pow(w, i, R) % R)
    back = orc.extended_to_coeff(ext, ext_k, fr_bytes(g))
    bi = bytes_to_ints(orc.from_mont(1, back))
    assert bi[:1 << k] == coeffs and all(v == 0 for v in bi[1 << k:])


def test_blake2b_matches_hashlib(orc):
    rng = random.Random(11)
    for ln in (0, 1, 63, 64, 65, 127, 128, 129, 255, 256, 257, 1000):
        msg = bytes(rng.randrange(256) for _ in range(ln))
        for pers in (b"Halo2-Transcript", b"Halo2-Verify-Key", None):
            kw = {"person": pers} if pers else {}
            assert orc.blake2b(msg, pers, 64) == hashlib.blake2b(msg, digest_size=64, **kw).digest()
        assert orc.blake2b(msg, None, 32) == hashlib.blake2b(msg, digest_size=32).digest()


def test_from_bytes_wide(orc):
    rng = random.Random(5)
    for _ in range(50):
        b = bytes(rng.randrange(256) for _ in range(64))
        got = bytes_to_ints(orc.from_mont(1, orc.fr_from_bytes_wide(b)))[0]
        assert got == int.from_bytes(b, "little") % R
    assert bytes_to_ints(orc.from_mont(1, orc.fr_from_bytes_wide(b"\xff" * 64)))[0] == ((1 << 512) - 1) % R


def test_transcript_matches_pymodel(orc):
    rng = random.Random(21)
    t, m = orc.Transcript(), pm.Blake2bTranscript()
    for step in range(40):
        kind = rng.randrange(3)
        if kind == 0:
            pt = pm.g1_mul(pm.G1, rng.randrange(1, R))
            t.common_point(np.frombuffer(pm.affine_bytes(pt), dtype=np.uint8)); m.common_point(pt)
        elif kind == 1:
            s = rng.randrange(R)
            t.common_scalar(fr_bytes(s)); m.common_scalar(s)
        else:
            assert bytes_to_ints(orc.from_mont(1, t.squeeze()))[0] == m.squeeze_challenge()
    assert bytes_to_ints(orc.from_mont(1, t.squeeze()))[0] == m.squeeze_challenge()


def test_gwc_accumulate_matches_pymodel(orc):
    rng = random.Random(77)
    k = 9
    omega = pm.omega_for(k)
    rots = [0, 0, -1, 1, 0, -6, 1, 0, 0, -1, 0]
    queries = [(pm.g1_mul(pm.G1, rng.randrange(1, R)), r, rng.randrange(R)) for r in rots]
    ws = [pm.g1_mul(pm.G1, rng.randrange(1, R)) for _ in range(len(set(rots)))]
    x, u, v = (rng.randrange(R) for _ in range(3))
    want = pm.gwc_accumulate(queries, ws, x, u, v, omega)
    got = orc.gwc_accumulate(
        np.frombuffer(b"".join(pm.affine_bytes(q[0]) for q in queries), dtype=np.uint8),
        [q[1] for q in queries],
        np.frombuffer(b"".join(pm.fr_mont_bytes(q[2]) for q in queries), dtype=np.uint8),
        np.frombuffer(b"".join(pm.affine_bytes(w) for w in ws), dtype=np.uint8),
        fr_bytes(x), fr_bytes(u), fr_bytes(v), fr_bytes(omega), gen_affine())
    for i in range(4):
        assert pm.affine_from_bytes(got[64 * i:64 * i + 64]) == want[i]
    with pytest.raises(ValueError):
        orc.gwc_accumulate(np.zeros(64, np.uint8), [0], np.zeros(32, np.uint8), np.zeros(128, np.uint8),
                           fr_bytes(1), fr_bytes(1), fr_bytes(1), fr_bytes(omega), gen_affine())


def test_fold_h(orc):
    rng = random.Random(13)
    hs = [pm.g1_mul(pm.G1, rng.randrange(1, R)) for _ in range(4)]
    xn = rng.randrange(R)
    want = pm.msm([pow(xn, i, R) for i in range(4)], hs)
    got = orc.fold_h(np.frombuffer(b"".join(pm.affine_bytes(h) for h in hs), dtype=np.uint8), fr_bytes(xn))
    assert pm.affine_from_bytes(got) == want


def test_point_compression_model():
    for k in (1, 2, 3, 7, 12345):
        pt = pm.g1_mul(pm.G1, k)
        assert pm.decompress_point(pm.compress_point(pt)) == pt
        assert pm.decompress_point(pm.compress_point(pm.g1_neg(pt))) == pm.g1_neg(pt)
    # 68-bit limbs of x(2G) (examples/simple-example.rs:535-537 convention)
    x2 = KAT[2][0]
    limbs = [(x2 >> (68 * i)) & ((1 << 68) - 1) for i in range(4)]
    assert limbs == [0x8d3c208c16d87cfd3, 0x85d97816a916871ca, 0xa029b85045b681815, 0x30644e72e131]
