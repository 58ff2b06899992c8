"""CPU tests of the mul_var witness oracle (oracle/mulvar.py, row f4): its cells satisfy the identities the circuit enforces
(a*b = q*p + r over the integers through 68-bit limb products, both CRT halves), its limb packing is the reference's
(examples/simple-example.rs:535-537; SURVEY App. A known answer), and its result is s*P of the curve model."""
import random

import numpy as np

from oracle import mulvar as mv
from oracle import pymodel as pm


def test_limb_packing_known_answer():
    x2g = 0x030644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd3     # x(2G), SURVEY App. A
    assert mv.limbs(x2g) == [0x8d3c208c16d87cfd3, 0x85d97816a916871ca, 0xa029b85045b681815, 0x30644e72e131]
    assert mv.NEG_P + mv.P == 1 << 272 and mv.LEN == 43508


def test_records_satisfy_the_circuit_identities():
    rng = random.Random(5)
    edge = [0, 1, 2, mv.P - 1, mv.P - 2, (1 << 68) - 1, ((1 << 68) - 1) << 68, (1 << 204) - 1, (1 << 253), mv.P >> 1]
    pairs = [(a, b) for a in edge for b in edge] + [(rng.randrange(mv.P), rng.randrange(mv.P)) for _ in range(300)]
    for a, b in pairs:
        cells, r = mv.record(a, b)
        assert len(cells) == mv.REC and r == a * b % mv.P
        assert mv.check_record(cells) == (a, b, r)
        assert max(c.bit_length() for c in cells) <= 141


def test_witness_of_one_mul_var():
    rng = random.Random(9)
    aux = pm.g1_mul(pm.G1, 0xabcdef123457)
    for scalar in (1, 2, pm.R - 1, rng.randrange(pm.R)):
        point = pm.g1_mul(pm.G1, rng.randrange(1, pm.R))
        q, cells, status = mv.mulvar_witness(point, scalar, aux)
        assert status == 0 and q == pm.g1_mul(point, scalar) and len(cells) == mv.LEN and None not in cells
        assert sum(c << i for i, c in enumerate(cells[:mv.BITS])) == scalar
        acc = aux
        for step in range(mv.BITS):
            base = mv.BITS + step * mv.STEP
            for j in range(7):
                mv.check_record(cells[base + mv.REC * j:base + mv.REC * (j + 1)])
            d = pm.g1_add(acc, acc)
            acc = pm.g1_add(d, point) if (scalar >> (mv.BITS - 1 - step)) & 1 else d
            assert cells[base + 7 * mv.REC:base + 7 * mv.REC + 8] == mv.limbs(d[0]) + mv.limbs(d[1])
            assert cells[base + 7 * mv.REC + 8:base + 7 * mv.REC + 16] == mv.limbs(acc[0]) + mv.limbs(acc[1])
        base = mv.BITS + mv.BITS * mv.STEP
        assert cells[base + 3 * mv.REC:base + 3 * mv.REC + 8] == mv.limbs(q[0]) + mv.limbs(q[1])
    # what the incomplete formulas cannot witness is reported, not mis-stated
    assert mv.mulvar_witness(pm.g1_mul(pm.G1, 5), 0, aux)[2] == 1 + mv.BITS          # s = 0: the final addition cancels
    assert mv.mulvar_witness(None, 7, aux)[2] == 0xffffffff                            # the identity has no affine cells
    assert mv.mulvar_witness(pm.g1_mul(aux, 2), 7, aux)[2] == 1                        # first addition meets P = 2 aux


def test_compiled_restatement_matches_the_big_integer_one(orc):
    """oracle/oracle.cpp `orc_mulvar_witness` (the compiled CPU figure of the benchmark, and the checker of large batches) writes the
    cells of oracle/mulvar.py, cell for cell, and reports the same status for what cannot be witnessed."""
    rng = random.Random(3)
    aux = pm.g1_mul(pm.G1, 0xabcdef123457)
    pts = [pm.g1_mul(pm.G1, rng.randrange(1, pm.R)) for _ in range(4)] + [None, pm.g1_mul(aux, 2)]
    sc = [1, pm.R - 1, rng.randrange(pm.R), 0, 7, 7]
    pb = np.frombuffer(b"".join(pm.affine_bytes(p) for p in pts), dtype=np.uint8)
    sb = np.frombuffer(b"".join(pm.fr_mont_bytes(s) for s in sc), dtype=np.uint8)
    ab = np.frombuffer(pm.affine_bytes(aux), dtype=np.uint8)
    assert orc.mulvar_witness_len() == mv.LEN
    res, cells, status = orc.mulvar_witness(pb, sb, ab)
    for i, (p, s) in enumerate(zip(pts, sc)):
        q, want, st = mv.mulvar_witness(p, s, aux)
        assert st == status[i]
        if st == 0:
            assert bytes(cells[32 * mv.LEN * i:32 * mv.LEN * (i + 1)]) == b"".join(pm.fr_mont_bytes(c) for c in want)
            assert pm.affine_from_bytes(bytes(res[64 * i:64 * i + 64])) == q
        else:
            assert bytes(res[64 * i:64 * i + 64]) == bytes(64)
    res2, none, status2 = orc.mulvar_witness(pb, sb, ab, want_cells=False, threads=2)
    assert none is None and bytes(res2) == bytes(res) and list(status2) == list(status)
