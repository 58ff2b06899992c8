"""GPU tests of row f4: the witness cells of batched non-native mul_var (csrc/mulvar.cu) against the big-integer oracle
(oracle/mulvar.py), cell for cell, and the products against best_multiexp-style scalar multiplication of the curve model."""
import random

import numpy as np
import pytest

import halo2_aggregation_b200 as h2a
from oracle import mulvar as mv
from oracle import pymodel as pm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = h2a.Context(0)
    yield c
    c.close()


def pts_bytes(points):
    return np.frombuffer(b"".join(pm.affine_bytes(p) for p in points), dtype=np.uint8)


def frs_bytes(vals):
    return np.frombuffer(b"".join(pm.fr_mont_bytes(v) for v in vals), dtype=np.uint8)


AUX = pm.g1_mul(pm.G1, 0xabcdef123457)


def test_witness_cells_match_oracle(ctx):
    rng = random.Random(21)
    assert ctx.mulvar_witness_len() == mv.LEN
    scalars = [1, 2, pm.R - 1, (1 << 253) + 1, rng.randrange(pm.R), rng.randrange(pm.R), rng.randrange(1 << 68)]
    points = [pm.G1, pm.g1_mul(pm.G1, 2)] + [pm.g1_mul(pm.G1, rng.randrange(1, pm.R)) for _ in range(len(scalars) - 2)]
    res, wit, status = ctx.mulvar_witness(pts_bytes(points), frs_bytes(scalars), pts_bytes([AUX]))
    assert not status.any()
    for i, (p, s) in enumerate(zip(points, scalars)):
        q, cells, st = mv.mulvar_witness(p, s, AUX)
        assert st == 0 and pm.affine_from_bytes(bytes(res[64 * i:64 * i + 64])) == q == pm.g1_mul(p, s)
        got = wit[32 * mv.LEN * i:32 * mv.LEN * (i + 1)]
        want = frs_bytes(cells)
        if bytes(got) != bytes(want):
            first = next(j for j in range(mv.LEN) if bytes(got[32 * j:32 * j + 32]) != bytes(want[32 * j:32 * j + 32]))
            raise AssertionError("entry %d: cell %d of %d differs" % (i, first, mv.LEN))


def test_batch_of_one_proof_worth_of_mul_var(ctx):
    """37 mul_var (one aggregated proof, src/multiopen.rs:393-492 + src/vanishing.rs:181-187) x 8 proofs in one launch:
    products against the curve model, records of a sample of entries against the circuit's identities."""
    rng = random.Random(22)
    m = 37 * 8
    base_pts = [pm.g1_mul(pm.G1, rng.randrange(1, pm.R)) for _ in range(16)]
    points = [base_pts[rng.randrange(16)] for _ in range(m)]
    scalars = [rng.randrange(1, pm.R) for _ in range(m)]
    res, wit, status = ctx.mulvar_witness(pts_bytes(points), frs_bytes(scalars), pts_bytes([AUX]))
    assert not status.any()
    for i in range(0, m, 7):
        assert pm.affine_from_bytes(bytes(res[64 * i:64 * i + 64])) == pm.g1_mul(points[i], scalars[i])
    for i in (0, 100, m - 1):
        cells = [pm.fr_from_mont_bytes(bytes(wit[32 * (mv.LEN * i + j):32 * (mv.LEN * i + j) + 32])) for j in range(mv.LEN)]
        for step in (0, 17, 253):
            base = mv.BITS + step * mv.STEP
            for j in range(7):
                mv.check_record(cells[base + mv.REC * j:base + mv.REC * (j + 1)])


def test_unwitnessable_entries_are_reported(ctx):
    points = [pm.g1_mul(pm.G1, 5), None, pm.g1_mul(AUX, 2), pm.g1_mul(pm.G1, 9)]
    scalars = [0, 7, 7, 3]
    with pytest.raises(h2a.H2AError) as e:
        ctx.mulvar_witness(pts_bytes(points), frs_bytes(scalars), pts_bytes([AUX]))
    assert list(e.value.status) == [1 + mv.BITS, 0xffffffff, 1, 0]
    with pytest.raises(h2a.H2AError):
        ctx.mulvar_witness(pts_bytes(points[3:]), frs_bytes(scalars[3:]), np.zeros(64, np.uint8))      # identity as the auxiliary point
    res, _, _ = ctx.mulvar_witness(pts_bytes(points[3:]), frs_bytes(scalars[3:]), pts_bytes([AUX]), want_witness=False)
    assert pm.affine_from_bytes(bytes(res)) == pm.g1_mul(pm.G1, 27)
