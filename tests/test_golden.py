"""The committed fixtures (tests/golden/oracle_vectors.json, written by tests/golden/make_golden.py) against the
oracle (CPU) and against the CUDA library through the C ABI (GPU)."""
import hashlib
import json
import os

import numpy as np
import pytest

import circuits
from oracle import plonk as pk
from oracle import pymodel as pm

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(HERE, "golden", "oracle_vectors.json")) as f:
        return json.load(f)


def test_oracle_reproduces_the_fixtures(orc, gold):
    for k, h in gold["g1_multiples_affine_mont_hex"].items():
        g = np.frombuffer(pm.affine_bytes(pm.G1), dtype=np.uint8)
        s = np.frombuffer(pm.fr_mont_bytes(int(k)), dtype=np.uint8)
        assert bytes(orc.g1_mul(g, s)).hex() == h
    for k, h in gold["fr_root_of_unity_mont_hex"].items():
        assert bytes(orc.fr_root_of_unity(int(k))).hex() == h
    assert bytes(orc.gen_bases(1, 4)).hex() == gold["gen_bases_seed1_first4_hex"]
    assert bytes(orc.gen_scalars(2, 4)).hex() == gold["gen_scalars_seed2_first4_hex"]
    n = 1 << 10
    assert bytes(orc.msm(orc.gen_bases(1, n), orc.gen_scalars(2, n))).hex() == gold["msm_2^10_seed1_seed2_affine_hex"]
    a = orc.gen_scalars(3, 1 << 8)
    assert hashlib.sha256(bytes(orc.fft(a, 8, orc.fr_root_of_unity(8)))).hexdigest() == gold["ntt_2^8_seed3_sha"]
    t = orc.Transcript()
    t.common_point(np.frombuffer(pm.affine_bytes(pm.G1), dtype=np.uint8))
    t.common_point(np.frombuffer(pm.affine_bytes(pm.g1_mul(pm.G1, 2)), dtype=np.uint8))
    t.common_scalar(np.frombuffer(pm.fr_mont_bytes(5), dtype=np.uint8))
    got = [hex(pm.fr_from_mont_bytes(t.squeeze())), hex(pm.fr_from_mont_bytes(t.squeeze()))]
    assert got == gold["transcript_challenges_hex"]


def test_oracle_proof_fixture(orc, gold):
    g = gold["my_circuit_k6"]
    c = circuits.my_circuit(k=6, table_bits=4)
    params, keys = circuits.setup(orc, c, s=int(g["setup_secret_hex"], 16), vk_hash=int(g["vk_hash_hex"], 16))
    proof, inst = pk.create_proof(orc, params, c["shape"], keys, c["instance"], c["advice"], seed=g["blind_seed"])
    assert proof.hex() == g["proof_hex"]
    assert pm.affine_bytes(inst[0]).hex() == g["instance_commitment_hex"]
    res = pk.verify_proof(c["shape"], keys.fixed_commitments, keys.sigma_commitments, keys.vk_hash, inst, bytes.fromhex(g["proof_hex"]))
    assert b"".join(pm.affine_bytes(res[k]) for k in ("e", "f", "w", "zw")).hex() == g["efwzw_hex"]
    assert {k: hex(res[k]) for k in g["challenges_hex"]} == g["challenges_hex"]


@pytest.mark.gpu
def test_library_reproduces_the_fixtures(gold):
    """No oracle on this path: fixtures in, C ABI out."""
    import halo2_aggregation_b200 as h2a
    ctx = h2a.Context(0)
    d = ctx.dev_alloc(64 * 4)
    ctx.gen_bases_dev(1, 4, d)
    assert bytes(ctx.d2h(d, 256)).hex() == gold["gen_bases_seed1_first4_hex"]
    ctx.gen_scalars_dev(2, 4, d)
    assert bytes(ctx.d2h(d, 128)).hex() == gold["gen_scalars_seed2_first4_hex"]
    ctx.dev_free(d)
    n = 1 << 10
    db, ds = ctx.dev_alloc(64 * n), ctx.dev_alloc(32 * n)
    ctx.gen_bases_dev(1, n, db); ctx.gen_scalars_dev(2, n, ds)
    hb = ctx.bases_from_device(db, n)
    assert bytes(ctx.msm_dev(hb, ds, n)).hex() == gold["msm_2^10_seed1_seed2_affine_hex"]
    hb.free()
    ctx.gen_scalars_dev(3, 1 << 8, ds)
    a = ctx.d2h(ds, 32 << 8)
    assert hashlib.sha256(bytes(ctx.ntt(a, 8, h2a.fr_root_of_unity(8)))).hexdigest() == gold["ntt_2^8_seed3_sha"]
    for k, h in gold["fr_root_of_unity_mont_hex"].items():
        assert bytes(h2a.fr_root_of_unity(int(k))).hex() == h
    # the proof fixture through the verifier glue
    g = gold["my_circuit_k6"]
    shape = circuits.my_circuit(k=6, table_bits=4)["shape"]
    circ = h2a.Circuit(ctx, shape, np.zeros(0, np.uint8))
    hexs = lambda xs: np.frombuffer(bytes.fromhex("".join(xs)), dtype=np.uint8)
    vk_hash = ctx.field_op(1, "to_mont", np.frombuffer(int(g["vk_hash_hex"], 16).to_bytes(32, "little"), dtype=np.uint8))
    circ.set_vk(hexs(g["fixed_commitments_hex"]), hexs(g["sigma_commitments_hex"]), vk_hash)
    got = circ.verify(hexs([g["instance_commitment_hex"]]), bytes.fromhex(g["proof_hex"]))
    assert bytes(got).hex() == g["efwzw_hex"]
    circ.free(); ctx.dev_free(db); ctx.dev_free(ds); ctx.close()
