"""GPU tests of the whole-proof glue: the library's verifier replays proofs written
by the oracle prover and must return the oracle's (e, f, w, zw) bit for bit; the library's prover must write
byte-identical proofs."""
import numpy as np
import pytest

import circuits
import halo2_aggregation_b200 as h2a
from oracle import plonk as pk
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
    return np.frombuffer(b"".join(pm.fr_mont_bytes(v) for v in vals), dtype=np.uint8) if vals else np.zeros(0, np.uint8)


def make_circuit(ctx, c, keys):
    shape = c["shape"]
    circ = h2a.Circuit(ctx, shape, frs_bytes(shape.constants))
    circ.set_vk(pts_bytes(keys.fixed_commitments), pts_bytes(keys.sigma_commitments), frs_bytes([keys.vk_hash]))
    return circ


def efwzw_bytes(res):
    return pts_bytes([res["e"], res["f"], res["w"], res["zw"]])


@pytest.mark.parametrize("which", ["my_circuit", "wide"])
def test_verifier_glue_matches_oracle(ctx, orc, which):
    c = circuits.my_circuit(k=6, table_bits=4) if which == "my_circuit" else circuits.wide_circuit(k=6)
    params, keys = circuits.setup(orc, c)
    proof, inst = pk.create_proof(orc, params, c["shape"], keys, c["instance"], c["advice"], seed=11)
    want = pk.verify_proof(c["shape"], keys.fixed_commitments, keys.sigma_commitments, keys.vk_hash, inst, proof)
    assert pk.pairing_relation_holds(want, params.s)
    circ = make_circuit(ctx, c, keys)
    got = circ.verify(pts_bytes(inst), proof)
    assert bytes(got) == bytes(efwzw_bytes(want))
    # malformed proofs are refused with H2A_ERR_PROOF, never crash
    for bad in (proof[:-1], proof + b"\0", proof[:40], b"\xff" * len(proof)):
        with pytest.raises(h2a.H2AError) as e:
            circ.verify(pts_bytes(inst), bad)
        assert e.value.code == -5
    # a flipped evaluation still parses but no longer satisfies the pairing relation
    t = bytearray(proof)
    t[32 * (c["shape"].num_advice + 2 * len(c["shape"].lookups) + 2 + len(c["shape"].lookups) + 1 + 4) + 1] ^= 1
    got_bad = circ.verify(pts_bytes(inst), bytes(t))
    want_bad = pk.verify_proof(c["shape"], keys.fixed_commitments, keys.sigma_commitments, keys.vk_hash, inst, bytes(t))
    assert bytes(got_bad) == bytes(efwzw_bytes(want_bad)) and not pk.pairing_relation_holds(want_bad, params.s)
    circ.free()


def test_verifier_glue_batch_of_64(ctx, orc):
    """BASELINE config 5: 64 proofs of the sample circuit, all 256 sums in one launch."""
    c = circuits.my_circuit(k=5, table_bits=3)
    params, keys = circuits.setup(orc, c)
    circ = make_circuit(ctx, c, keys)
    proofs, insts, wants = [], [], []
    for s in range(8):
        ci = circuits.my_circuit(k=5, table_bits=3, a=1 + s % 7, b=2 + s % 5, seed=s)   # same fixed columns, other witnesses
        proof, inst = pk.create_proof(orc, params, ci["shape"], keys, ci["instance"], ci["advice"], seed=s)
        proofs.append(proof); insts.append(inst[0])
        wants.append(pk.verify_proof(ci["shape"], keys.fixed_commitments, keys.sigma_commitments, keys.vk_hash, inst, proof))
    batch = [proofs[i % 8] for i in range(64)]
    binst = pts_bytes([insts[i % 8] for i in range(64)])
    got = circ.verify_batch(binst, batch)
    for i in range(64):
        assert bytes(got[i]) == bytes(efwzw_bytes(wants[i % 8])), i
    assert all(pk.pairing_relation_holds(w, params.s) for w in wants)
    circ.free()


def cols_bytes(cols):
    return np.frombuffer(b"".join(pm.fr_mont_bytes(v) for col in cols for v in col), dtype=np.uint8)


def prove_with_library(ctx, orc, c, params, keys, seed):
    shape = c["shape"]
    g, gl = ctx.upload_bases(params.g), ctx.upload_bases(params.g_lagrange)
    circ = h2a.Circuit(ctx, shape, frs_bytes(shape.constants))
    circ.set_keys(g, gl, cols_bytes(c["fixed"]), cols_bytes(keys.sigmas), frs_bytes([keys.vk_hash]), frs_bytes([shape.coset_shift]))
    proof, inst = circ.prove(cols_bytes(c["instance"]), cols_bytes(c["advice"]), frs_bytes(pk.blinds_buffer(shape, seed)))
    return circ, proof, inst, (g, gl)


@pytest.mark.parametrize("which,k", [("my_circuit", 6), ("wide", 6), ("my_circuit", 9), ("many_rotations", 6)])
def test_prover_writes_byte_identical_proofs(ctx, orc, which, k):
    """Row a1: same transcript, same commitments, same evaluations, same witnesses -> identical bytes
    (k = 9 is the reference's own sample size, examples/simple-example.rs:561; `many_rotations` opens at twelve points)."""
    c = {"my_circuit": lambda: circuits.my_circuit(k=k, table_bits=min(8, k - 2)), "wide": lambda: circuits.wide_circuit(k=k),
         "many_rotations": lambda: circuits.many_rotations_circuit(k=k)}[which]()
    params, keys = circuits.setup(orc, c)
    want, want_inst = pk.create_proof(orc, params, c["shape"], keys, c["instance"], c["advice"], seed=5)
    circ, proof, inst, handles = prove_with_library(ctx, orc, c, params, keys, seed=5)
    fc, sc = circ.get_vk(c["shape"].num_fixed, len(c["shape"].perm_columns))
    assert bytes(fc) == bytes(pts_bytes(keys.fixed_commitments)) and bytes(sc) == bytes(pts_bytes(keys.sigma_commitments))
    assert bytes(inst) == bytes(pts_bytes(want_inst))
    assert len(proof) == len(want) == circ.proof_len()
    if proof != want:
        first = next(i for i in range(0, len(want), 32) if proof[i:i + 32] != want[i:i + 32])
        raise AssertionError("proofs differ first at 32-byte item %d of %d" % (first // 32, len(want) // 32))
    # and the library's own verifier glue accepts what it wrote
    res = pk.verify_proof(c["shape"], keys.fixed_commitments, keys.sigma_commitments, keys.vk_hash, want_inst, proof)
    assert bytes(circ.verify(inst, proof)) == bytes(efwzw_bytes(res)) and pk.pairing_relation_holds(res, params.s)
    # a lookup input outside the table is reported, not silently proven
    bad = [col[:] for col in c["advice"]]
    bad[0][0] = 12345678
    with pytest.raises(h2a.H2AError):
        circ.prove(cols_bytes(c["instance"]), cols_bytes(bad), frs_bytes(pk.blinds_buffer(c["shape"], 5)))
    circ.free()
    for h in handles:
        h.free()


def test_kzg_setup_matches_oracle(ctx, orc):
    """Row a10: Setup::new for a given secret — g and g_lagrange equal the oracle's, point for point."""
    s = 0x1234567890abcdef1234567890abcdef
    for k in (1, 6):
        params = pk.Params(orc, k, s)
        g, gl = ctx.kzg_setup(k, frs_bytes([s]))
        assert bytes(g.download()) == bytes(params.g)
        assert bytes(gl.download()) == bytes(params.g_lagrange)
        g.free(); gl.free()
    # commit_lagrange of the constant-one column is [1] G: sum of the Lagrange basis
    g, gl = ctx.kzg_setup(10, frs_bytes([s]))
    ones = frs_bytes([1] * 1024)
    assert pm.affine_from_bytes(ctx.msm(gl, ones)) == pm.G1
    g.free(); gl.free()


def test_prover_rejects_what_the_verifier_would(ctx, orc):
    """An identity commitment (an all-zero, unblinded advice column) cannot be absorbed by the transcript — the verifier's
    read_point rejects it — so create_proof fails instead of writing a proof no one can check; short column buffers and
    a shape whose permutation entry does not name its own rotation-0 query are refused at the boundary."""
    c = circuits.my_circuit(k=6, table_bits=4)
    params, keys = circuits.setup(orc, c)
    circ, proof, inst, handles = prove_with_library(ctx, orc, c, params, keys, seed=5)
    n = 1 << 6
    zero_adv = [c["advice"][0], [0] * n]
    with pytest.raises(h2a.H2AError, match="identity"):
        circ.prove(cols_bytes(c["instance"]), cols_bytes(zero_adv), frs_bytes(pk.blinds_buffer(c["shape"], 5)))
    # the failed call left nothing in flight: the same circuit proves again, byte for byte
    again, _ = circ.prove(cols_bytes(c["instance"]), cols_bytes(c["advice"]), frs_bytes(pk.blinds_buffer(c["shape"], 5)))
    assert again == proof
    with pytest.raises(h2a.H2AError):
        circ.prove(cols_bytes(c["instance"]), cols_bytes(c["advice"])[:-32], frs_bytes(pk.blinds_buffer(c["shape"], 5)))
    circ.free()
    for h in handles:
        h.free()
    sh = c["shape"]
    bad = pk.Shape(k=sh.k, blinding_factors=sh.bf, degree=sh.degree, num_instance=1, num_advice=2, num_fixed=4,
                   advice_queries=sh.advice_queries, fixed_queries=sh.fixed_queries, instance_queries=sh.instance_queries, gates=sh.gates,
                   constants=[], lookups=sh.lookups, perm_columns=[(circuits.I, 0, 0), (circuits.F, 0, 3), (circuits.A, 0, 1), (circuits.A, 1, 1)])
    with pytest.raises(h2a.H2AError):
        h2a.Circuit(ctx, bad, frs_bytes([]))


@pytest.mark.parametrize("k", [11, 14])
def test_large_proof_verifies(ctx, orc, k):
    """Beyond the oracle prover's reach: prove on the device with device-generated parameters, then check the proof
    with the oracle verifier and the pairing relation (multi-pass NTT, recursive scans, chunked Kate division)."""
    s = 0x0badc0ffee0ddf00d1234567890abcdef
    c = circuits.my_circuit(k=k, table_bits=8, a=251, b=97, constant=4242)
    shape = c["shape"]
    g, gl = ctx.kzg_setup(k, frs_bytes([s]))
    sigmas = pk.build_sigmas(shape, c["cycles"])
    circ = h2a.Circuit(ctx, shape, frs_bytes(shape.constants))
    circ.set_keys(g, gl, cols_bytes(c["fixed"]), cols_bytes(sigmas), frs_bytes([77]), frs_bytes([shape.coset_shift]))
    proof, inst = circ.prove(cols_bytes(c["instance"]), cols_bytes(c["advice"]), frs_bytes(pk.blinds_buffer(shape, 9)))
    fc, sc = circ.get_vk(shape.num_fixed, len(shape.perm_columns))
    to_pts = lambda b: [pm.affine_from_bytes(b[64 * i:64 * i + 64]) for i in range(len(b) // 64)]
    res = pk.verify_proof(shape, to_pts(fc), to_pts(sc), 77, to_pts(inst), proof)
    assert pk.pairing_relation_holds(res, s)
    assert bytes(circ.verify(inst, proof)) == bytes(efwzw_bytes(res))
    phases = circ.prove_phases()
    assert len(phases) >= 8 and all(ms >= 0 for _, ms in phases)
    # a wrong witness (broken product) still yields a proof, which must NOT verify
    bad = [col[:] for col in c["advice"]]
    bad[0][4] = (bad[0][4] + 1) % pm.R
    proof2, inst2 = circ.prove(cols_bytes(c["instance"]), cols_bytes(bad), frs_bytes(pk.blinds_buffer(shape, 9)))
    res2 = pk.verify_proof(shape, to_pts(fc), to_pts(sc), 77, to_pts(inst2), proof2)
    assert not pk.pairing_relation_holds(res2, s)
    circ.free(); g.free(); gl.free()


def test_keygen_sigmas_from_copy_constraints(ctx, orc):
    """Row a11: the permutation assembly's sigma columns are delta^col' * omega^row' of its own mapping, and a proof
    made with them verifies (the copy constraints of the sample circuit hold in its witness)."""
    k = 7
    c = circuits.my_circuit(k=k, table_bits=4)
    shape = c["shape"]
    m = len(shape.perm_columns)
    asm = h2a.PermutationAssembly(m, k)
    for cyc in c["cycles"]:
        for (ca, ra), (cb, rb) in zip(cyc, cyc[1:]):
            asm.copy(ca, ra, cb, rb)
    nxt = asm.mapping()
    got = asm.sigmas(ctx, frs_bytes([shape.omega]), frs_bytes([pk.DELTA]))
    n = 1 << k
    om = [pow(shape.omega, i, pm.R) for i in range(n)]
    want = [pow(pk.DELTA, int(v) >> k, pm.R) * om[int(v) & (n - 1)] % pm.R for v in nxt]
    assert bytes(got) == bytes(frs_bytes(want))
    # the same equivalence classes as the oracle's cycle list, so both sigma sets accept the same witnesses
    s = 0x5eed5eed5eed
    g, gl = ctx.kzg_setup(k, frs_bytes([s]))
    circ = h2a.Circuit(ctx, shape, frs_bytes(shape.constants))
    circ.set_keys(g, gl, cols_bytes(c["fixed"]), got, frs_bytes([5]), frs_bytes([shape.coset_shift]))
    proof, inst = circ.prove(cols_bytes(c["instance"]), cols_bytes(c["advice"]), frs_bytes(pk.blinds_buffer(shape, 3)))
    fc, sc = circ.get_vk(shape.num_fixed, m)
    to_pts = lambda b: [pm.affine_from_bytes(b[64 * i:64 * i + 64]) for i in range(len(b) // 64)]
    res = pk.verify_proof(shape, to_pts(fc), to_pts(sc), 5, to_pts(inst), proof)
    assert pk.pairing_relation_holds(res, s)
    circ.free(); g.free(); gl.free()
