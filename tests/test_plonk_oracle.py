"""The oracle PLONK+KZG prover/verifier (oracle/plonk.py) is self-consistent: proofs it writes are read
back in the reference's wire order, the accumulated (e, f, w, zw) satisfy the final pairing equation (checked
as a discrete-log relation with the known setup secret), and tampering breaks it.  CPU only."""
import pytest

import circuits
from oracle import plonk as pk
from oracle import pymodel as pm


@pytest.fixture(scope="module")
def small(orc):
    c = circuits.my_circuit(k=6, table_bits=4)
    params, keys = circuits.setup(orc, c)
    proof, inst = pk.create_proof(orc, params, c["shape"], keys, c["instance"], c["advice"], seed=7)
    return c, params, keys, proof, inst


def test_my_circuit_proof_verifies(small):
    c, params, keys, proof, inst = small
    res = pk.verify_proof(c["shape"], keys.fixed_commitments, keys.sigma_commitments, keys.vk_hash, inst, proof)
    assert pk.pairing_relation_holds(res, params.s)
    # proof size: 2 advice + 2 lookup perm + 2 perm z + 1 lookup z + 1 random + 4 h + 4 W points,
    # (1+3+4) query evals + 1 random + 4 sigma + (3+2) perm + 5 lookup = 23 scalars  (SURVEY: "~1.2 kB")
    assert len(proof) == 32 * (16 + 23)
    rots = sorted({q[1] for q in res["queries"]})
    assert rots == [-6, -1, 0, 1] and len(res["ws"]) == 4
    assert len(res["queries"]) == 1 + 3 + (2 * 2 + 1) + 5 + 4 + 4 + 2


def test_proof_is_reproducible_and_seed_dependent(orc, small):
    c, params, keys, proof, _ = small
    again, _ = pk.create_proof(orc, params, c["shape"], keys, c["instance"], c["advice"], seed=7)
    other, _ = pk.create_proof(orc, params, c["shape"], keys, c["instance"], c["advice"], seed=8)
    assert again == proof and other != proof


def test_tampered_proof_fails(small):
    c, params, keys, proof, inst = small
    bad = bytearray(proof)
    bad[32 * 16 + 3] ^= 1       # first evaluation scalar
    res = pk.verify_proof(c["shape"], keys.fixed_commitments, keys.sigma_commitments, keys.vk_hash, inst, bytes(bad))
    assert not pk.pairing_relation_holds(res, params.s)
    wrong_inst = [pm.g1_add(inst[0], pm.G1)]
    res = pk.verify_proof(c["shape"], keys.fixed_commitments, keys.sigma_commitments, keys.vk_hash, wrong_inst, proof)
    assert not pk.pairing_relation_holds(res, params.s)


@pytest.mark.parametrize("cell", [(0, 4), (1, 3)])
def test_unsatisfied_witness_does_not_verify(orc, cell):
    """(0,4): a*b != out breaks the gate and a copy constraint; (1,3): rhs != b breaks a copy constraint and the gate."""
    c = circuits.my_circuit(k=6, table_bits=4)
    params, keys = circuits.setup(orc, c)
    c["advice"][cell[0]][cell[1]] += 1
    proof, inst = pk.create_proof(orc, params, c["shape"], keys, c["instance"], c["advice"], seed=7)
    res = pk.verify_proof(c["shape"], keys.fixed_commitments, keys.sigma_commitments, keys.vk_hash, inst, proof)
    assert not pk.pairing_relation_holds(res, params.s)


def test_lookup_violation_is_caught_by_the_prover(orc):
    c = circuits.my_circuit(k=6, table_bits=4, a=17)      # 17 is not in the 4-bit table
    params, keys = circuits.setup(orc, c)
    with pytest.raises(AssertionError):
        pk.create_proof(orc, params, c["shape"], keys, c["instance"], c["advice"], seed=7)


def test_wide_circuit_three_permutation_chunks(orc):
    c = circuits.wide_circuit(k=6)
    params, keys = circuits.setup(orc, c)
    proof, inst = pk.create_proof(orc, params, c["shape"], keys, c["instance"], c["advice"], seed=3)
    res = pk.verify_proof(c["shape"], keys.fixed_commitments, keys.sigma_commitments, keys.vk_hash, inst, proof)
    assert pk.pairing_relation_holds(res, params.s)
    assert sorted({q[1] for q in res["queries"]}) == [-7, -1, 0, 1]


def test_reference_workload_k9(orc):
    """The reference's own sample circuit size (examples/simple-example.rs:561: k = 9, u8 table)."""
    c = circuits.my_circuit(k=9, table_bits=8, a=200, b=13, constant=99)
    params, keys = circuits.setup(orc, c)
    proof, inst = pk.create_proof(orc, params, c["shape"], keys, c["instance"], c["advice"], seed=1)
    res = pk.verify_proof(c["shape"], keys.fixed_commitments, keys.sigma_commitments, keys.vk_hash, inst, proof)
    assert pk.pairing_relation_holds(res, params.s)
    assert inst[0] == pm.msm(c["public_inputs"], [pm.affine_from_bytes(params.g_lagrange[:64])])
