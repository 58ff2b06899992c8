#!/usr/bin/env python
"""Regenerates tests/golden/*.json from the oracle (oracle/ + tests/circuits.py).

The reference holds no golden vectors of its own (src/lib.rs:43-44 is an empty test module) and cannot be run here
(Rust, un-vendored dependencies), so these fixtures are OUR oracle's outputs: they pin the wire format, the
transcript, the synthetic-input generators and the known answers against accidental change, and they travel to the
GPU box where /root/reference does not exist.  Run from the repository root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import circuits  # noqa: E402
from oracle import loader as orc  # noqa: E402
from oracle import plonk as pk  # noqa: E402
from oracle import pymodel as pm  # noqa: E402


def main():
    out = {}
    # known answers in the library's wire layout
    out["g1_multiples_affine_mont_hex"] = {str(k): pm.affine_bytes(pm.g1_mul(pm.G1, k)).hex() for k in (1, 2, 3, 7, 30)}
    out["fr_root_of_unity_mont_hex"] = {str(k): bytes(orc.fr_root_of_unity(k)).hex() for k in (1, 9, 20, 23, 28)}
    # synthetic generators: first elements of the streams used by bench.py (seeds 1 and 2)
    out["gen_bases_seed1_first4_hex"] = bytes(orc.gen_bases(1, 4)).hex()
    out["gen_scalars_seed2_first4_hex"] = bytes(orc.gen_scalars(2, 4)).hex()
    n = 1 << 10
    out["msm_2^10_seed1_seed2_affine_hex"] = bytes(orc.msm(orc.gen_bases(1, n), orc.gen_scalars(2, n))).hex()
    a = orc.gen_scalars(3, 1 << 8)
    out["ntt_2^8_seed3_sha"] = __import__("hashlib").sha256(bytes(orc.fft(a, 8, orc.fr_root_of_unity(8)))).hexdigest()
    # transcript: absorb G, 2G, scalar 5, squeeze twice
    t = pm.Blake2bTranscript()
    t.common_point(pm.G1); t.common_point(pm.g1_mul(pm.G1, 2)); t.common_scalar(5)
    out["transcript_challenges_hex"] = [hex(t.squeeze_challenge()), hex(t.squeeze_challenge())]
    # a whole proof of the reference's sample circuit shape (small table), and what the verifier derives from it
    c = circuits.my_circuit(k=6, table_bits=4)
    params, keys = circuits.setup(orc, c)
    proof, inst = pk.create_proof(orc, params, c["shape"], keys, c["instance"], c["advice"], seed=7)
    res = pk.verify_proof(c["shape"], keys.fixed_commitments, keys.sigma_commitments, keys.vk_hash, inst, proof)
    assert pk.pairing_relation_holds(res, params.s)
    out["my_circuit_k6"] = {
        "setup_secret_hex": hex(params.s), "vk_hash_hex": hex(keys.vk_hash), "blind_seed": 7,
        "proof_hex": proof.hex(),
        "instance_commitment_hex": pm.affine_bytes(inst[0]).hex(),
        "efwzw_hex": b"".join(pm.affine_bytes(res[k]) for k in ("e", "f", "w", "zw")).hex(),
        "challenges_hex": {k: hex(res[k]) for k in ("theta", "beta", "gamma", "y", "x", "v", "u")},
        "fixed_commitments_hex": [pm.affine_bytes(p).hex() for p in keys.fixed_commitments],
        "sigma_commitments_hex": [pm.affine_bytes(p).hex() for p in keys.sigma_commitments],
    }
    with open(os.path.join(HERE, "oracle_vectors.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", os.path.join(HERE, "oracle_vectors.json"))


if __name__ == "__main__":
    main()
