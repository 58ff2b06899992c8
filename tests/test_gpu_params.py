"""GPU tests of row f3: parameter files — Params::write / Params::read (examples/simple-example.rs:679-691) — and
the verifier's view of the parameters (verifier_params, :693; commit_lagrange(public_inputs), :638-640)."""
import os

import numpy as np
import pytest

import halo2_aggregation_b200 as h2a
from oracle import plonk as pk
from oracle import pymodel as pm

pytestmark = pytest.mark.gpu

S = 0x1234567890abcdef1234567890abcdef


@pytest.fixture(scope="module")
def ctx():
    c = h2a.Context(0)
    yield c
    c.close()


def frs_bytes(vals):
    return np.frombuffer(b"".join(pm.fr_mont_bytes(v) for v in vals), dtype=np.uint8)


@pytest.mark.parametrize("compressed", [False, True])
def test_params_file_round_trip(ctx, orc, tmp_path, compressed):
    """write -> read returns the oracle's setup point for point, in both encodings; the trailer comes back unchanged."""
    k = 6
    params = pk.Params(orc, k, S)
    g, gl = ctx.kzg_setup(k, frs_bytes([S]))
    path = str(tmp_path / "halo2-6.params")
    trailer = np.arange(128, dtype=np.uint8)
    ctx.params_write(path, k, g, gl, compressed=compressed, trailer=trailer)
    assert os.path.getsize(path) == 64 + 2 * (1 << k) * (32 if compressed else 64) + 128 + 64
    k2, g2, gl2, tr = ctx.params_read(path)
    assert k2 == k and bytes(tr) == bytes(trailer)
    assert bytes(g2.download()) == bytes(params.g)
    assert bytes(gl2.download()) == bytes(params.g_lagrange)
    if compressed:   # the file holds the proof's point encoding of the oracle's points
        raw = open(path, "rb").read()
        pts = [pm.affine_from_bytes(bytes(params.g[64 * i:64 * i + 64])) for i in range(4)]
        assert raw[64:64 + 128] == b"".join(pm.compress_point(p) for p in pts)
    for h in (g, gl, g2, gl2):
        h.free()


def test_params_file_spans_several_pieces(ctx, tmp_path):
    """2^19 points = 32 MiB per array: two staging pieces each, written and read through both pinned buffers; no trailer."""
    k = 19
    g, gl = ctx.kzg_setup(k, frs_bytes([S]))
    for compressed in (False, True):
        path = str(tmp_path / ("p%d.params" % compressed))
        ctx.params_write(path, k, g, gl, compressed=compressed)
        k2, g2, gl2, tr = ctx.params_read(path)
        assert k2 == k and tr is None
        assert np.array_equal(g2.download(), g.download())
        assert np.array_equal(gl2.download(), gl.download())
        g2.free(); gl2.free()
        os.remove(path)
    g.free(); gl.free()


def test_params_read_rejects_damage(ctx, tmp_path):
    k = 5
    g, gl = ctx.kzg_setup(k, frs_bytes([S]))
    path = str(tmp_path / "a.params")
    ctx.params_write(path, k, g, gl)
    raw = bytearray(open(path, "rb").read())

    def expect_fail(data, match):
        bad = str(tmp_path / "bad.params")
        open(bad, "wb").write(bytes(data))
        with pytest.raises(h2a.H2AError, match=match):
            ctx.params_read(bad)

    t = bytearray(raw); t[64 + 64 * 3 + 5] ^= 1
    expect_fail(t, "digest")                       # a flipped coordinate bit
    expect_fail(raw[:-1], "truncated")
    expect_fail(raw + b"\0", "trailing")
    expect_fail(b"NOTPARAM" + raw[8:], "magic")
    t = bytearray(raw); t[12] = 7
    expect_fail(t, "header|truncated")             # k no longer matches n
    expect_fail(raw[:40], "header")
    with pytest.raises(h2a.H2AError, match="cannot open"):
        ctx.params_read(str(tmp_path / "missing.params"))
    g.free(); gl.free()


def test_params_read_checks_curve_membership(ctx, orc, tmp_path):
    """A file with a valid digest whose point is not on the curve (written from a forged handle) is refused on load."""
    k = 5
    g, gl = ctx.kzg_setup(k, frs_bytes([S]))
    pts = g.download().copy()
    pts[64 * 7 + 32] ^= 1                             # y of point 7 off by one bit
    forged = ctx.upload_bases(pts)
    for compressed, what in ((False, r"g\[7\]"), ):
        path = str(tmp_path / "forged.params")
        ctx.params_write(path, k, forged, gl, compressed=compressed)
        with pytest.raises(h2a.H2AError, match=what):
            ctx.params_read(path)
    # compressed: an x with no square root of x^3 + 3
    x = 0
    while pow(x ** 3 + 3, (pm.P - 1) // 2, pm.P) == 1 or x == 0:
        x += 1
    path = str(tmp_path / "nosqrt.params")
    ctx.params_write(path, k, g, gl, compressed=True)
    raw = bytearray(open(path, "rb").read())
    raw[64 + 32 * 2:64 + 32 * 3] = x.to_bytes(32, "little")
    import hashlib
    raw[-64:] = hashlib.blake2b(bytes(raw[:-64]), digest_size=64, person=b"H2A-Params-File\0").digest()
    open(path, "wb").write(bytes(raw))
    with pytest.raises(h2a.H2AError, match=r"g\[2\]"):
        ctx.params_read(path)
    for h in (g, gl, forged):
        h.free()


def test_verifier_params_commit_lagrange(ctx, orc):
    """params_verifier.commit_lagrange(public_inputs): an MSM over the first public_inputs_size Lagrange bases, equal to the
    oracle's best_multiexp over the same points; a longer input is refused."""
    k, npi = 9, 40                                    # 5 points x 8 limbs, examples/simple-example.rs:668-672
    params = pk.Params(orc, k, S)
    g, gl = ctx.kzg_setup(k, frs_bytes([S]))
    view = ctx.verifier_params(gl, npi)
    assert len(view) == npi
    rng = np.random.default_rng(3)
    pis = [int.from_bytes(rng.bytes(9), "little") for _ in range(npi)]     # 68-bit limbs
    sc = frs_bytes(pis)
    want = orc.msm(np.frombuffer(bytes(params.g_lagrange[:64 * npi]), np.uint8), sc)
    assert bytes(ctx.msm(view, sc)) == bytes(want)
    with pytest.raises(h2a.H2AError):
        ctx.msm(view, frs_bytes(pis + [1]))
    with pytest.raises(h2a.H2AError):
        ctx.verifier_params(gl, (1 << k) + 1)
    view.free()
    assert bytes(ctx.msm(gl, sc)) == bytes(want)      # freeing the view left the parameters alone
    g.free(); gl.free()
