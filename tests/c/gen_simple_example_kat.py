#!/usr/bin/env python
"""Writes examples/simple_example_kat.h: the known answers of examples/simple_example.cpp — the reference's
examples/simple-example.rs (`MyCircuit` at k = 9, constant 7, a 2, b 3; :550-643) driven through include/h2agg.hpp — computed
with the oracle (oracle/plonk.py prover and verifier, oracle/pymodel.py big-integer model, oracle/mulvar.py):

  the KZG secret of the reference's XorShift seed, the verifying key's commitments, the proof bytes, the instance commitment,
  (e, f, w, zw), the 40 public inputs of the aggregation circuit (68-bit limbs, :535-548, :668-672) and their commitment over
  the verifier's parameters, check sums of the mul_var witness cells of s * W, and host-side answers (field helpers, the
  permutation the copy constraints give, a transcript challenge).

The example draws its randomness (blinding rows, the prover's blinds) from the reference's XorShift construction with a
(domain, index) pair folded into the seed; `derived_scalar` below is the same rule."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import circuits
from oracle import loader as orc
from oracle import mulvar as mv
from oracle import plonk as pk
from oracle import pymodel as pm

OUT = os.path.join(ROOT, "examples", "simple_example_kat.h")
K, AGG_K = 9, 10
VK_HASH = 0xC0FFEE
DOMAIN_ADVICE_BLINDING = 1


def derived_scalar(domain, index):
    seed = bytearray(pm.REFERENCE_SETUP_SEED)
    for i, b in enumerate(domain.to_bytes(4, "little") + index.to_bytes(4, "little")):
        seed[i] ^= b
    seed[8] ^= 0xA5
    return pm.setup_secret_from_seed(bytes(seed))


def assembly_mapping(n_cols, k, copies):
    """The permutation `keygen` builds from copy constraints (halo2 `permutation::keygen::Assembly::copy`): cells col * n + row,
    the smaller cycle is renamed into the larger, the two successors are exchanged."""
    n = 1 << k
    mapping = list(range(n_cols * n))
    aux = list(range(n_cols * n))
    sizes = [1] * (n_cols * n)
    for (ca, ra), (cb, rb) in copies:
        left, right = ca * n + ra, cb * n + rb
        if aux[left] == aux[right]:
            continue
        if sizes[aux[left]] < sizes[aux[right]]:
            left, right = right, left
        name = aux[left]
        sizes[name] += sizes[aux[right]]
        cell = right
        while aux[cell] != name:
            aux[cell] = name
            cell = mapping[cell]
        mapping[left], mapping[right] = mapping[right], mapping[left]
    return mapping


def carr(name, data):
    rows = [", ".join("0x%02x" % b for b in data[i:i + 32]) for i in range(0, len(data), 32)]
    return "static const unsigned char %s[%d] = {\n    %s};\n" % (name, len(data), ",\n    ".join(rows))


def pts(points):
    return b"".join(pm.affine_bytes(p) for p in points)


def frs(vals):
    return b"".join(pm.fr_mont_bytes(v) for v in vals)


def limbs68(v):
    return [(v >> (68 * i)) & ((1 << 68) - 1) for i in range(4)]


def build():
    c = circuits.my_circuit(k=K, table_bits=8, a=2, b=3, constant=7)
    shape = c["shape"]
    n, bf = shape.n, shape.bf
    for ci, col in enumerate(c["advice"]):
        for j in range(bf):
            col[n - bf + j] = derived_scalar(DOMAIN_ADVICE_BLINDING, ci * bf + j)
    secret = pm.setup_secret_from_seed(pm.REFERENCE_SETUP_SEED)
    params = pk.Params(orc, K, secret)
    copies = [(a, b) for cyc in c["cycles"] for a, b in zip(cyc, cyc[1:])]
    mapping = assembly_mapping(len(shape.perm_columns), K, copies)
    om = [pow(shape.omega, i, pm.R) for i in range(n)]
    sigmas = [[pow(pk.DELTA, mapping[j * n + row] >> K, pm.R) * om[mapping[j * n + row] & (n - 1)] % pm.R for row in range(n)]
              for j in range(len(shape.perm_columns))]
    keys = pk.Keys(orc, params, shape, c["fixed"], sigmas, VK_HASH)
    real_blind = pk.blind
    pk.blind = lambda seed, obj, i: derived_scalar(obj, i)      # the example's randomness instead of the oracle's own stream
    try:
        proof, inst = pk.create_proof(orc, params, shape, keys, c["instance"], c["advice"], seed=0)
    finally:
        pk.blind = real_blind
    res = pk.verify_proof(shape, keys.fixed_commitments, keys.sigma_commitments, keys.vk_hash, inst, proof)
    assert pk.pairing_relation_holds(res, secret)
    quad = [res["e"], res["f"], res["w"], res["zw"]]
    # public inputs of the aggregation circuit: limbs of the instance commitment, then of e, f, w, zw (:668-672)
    public_inputs = []
    for p in [inst[0]] + quad:
        public_inputs += limbs68(p[0]) + limbs68(p[1])
    agg = pk.Params(orc, AGG_K, secret)
    pi_commitment = pk.commit(orc, agg.g_lagrange[:64 * len(public_inputs)], public_inputs)
    # mul_var(W, s) from the auxiliary point g[5]
    aux = pm.affine_from_bytes(bytes(params.g[64 * 5:64 * 6]))
    q, cells, st = mv.mulvar_witness(res["w"], secret, aux)
    assert st == 0 and q == pm.g1_add(pm.g1_add(res["zw"], res["f"]), res["e"]) and len(cells) == mv.LEN
    plain = sum(cells) % pm.R
    weighted = sum((i + 1) * v for i, v in enumerate(cells)) % pm.R
    # host-side answers
    tr = pm.Blake2bTranscript()
    tr.common_scalar(VK_HASH)
    tr.common_point(pm.G1)
    challenge = tr.squeeze_challenge()
    moved = [(i, v) for i, v in enumerate(mapping) if i != v]

    out = ["/* GENERATED by tests/c/gen_simple_example_kat.py from the oracle (oracle/plonk.py, oracle/pymodel.py, oracle/mulvar.py) — do not edit */\n"]
    out.append("#define KAT_K %d\n#define KAT_AGG_K %d\n#define KAT_MULVAR_LEN %d\n" % (K, AGG_K, mv.LEN))
    out.append(carr("KAT_SECRET", frs([secret])))
    out.append(carr("KAT_FR_SEVEN", frs([7])))
    out.append(carr("KAT_FR_MINUS_ONE", frs([pm.R - 1])))
    out.append(carr("KAT_FR_DELTA", frs([pk.DELTA])))
    out.append(carr("KAT_FR_C", frs([c["public_inputs"][0]])))
    out.append(carr("KAT_CHALLENGE", frs([challenge])))
    # a string of the shape `format!("{:?}", vk.pinned())` takes (longer than one 128-byte Blake2b block; no quotes to escape in C)
    sample = "PinnedVerificationKey { base_modulus: 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47, domain: PinnedEvaluationDomain { k: 9, extended_k: 11 } } " * 3
    out.append('static const char KAT_PINNED_SAMPLE[] = "%s";\n' % sample)
    out.append(carr("KAT_VK_HASH_SAMPLE", frs([pm.vk_hash_from_pinned(sample)])))
    out.append(carr("KAT_VK_HASH_EMPTY", frs([pm.vk_hash_from_pinned("")])))
    out.append("static const unsigned KAT_MAPPING_MOVED[%d][2] = {%s};\n" % (len(moved), ", ".join("{%d, %d}" % m for m in moved)))
    out.append(carr("KAT_G1", pts([pm.G1])))
    out.append(carr("KAT_TWO_G", pts([pm.g1_mul(pm.G1, 2)])))
    out.append(carr("KAT_TWO_G_LIMBS", frs(limbs68(pm.g1_mul(pm.G1, 2)[0]) + limbs68(pm.g1_mul(pm.G1, 2)[1]))))
    out.append(carr("KAT_FIXED_COMMITMENTS", pts(keys.fixed_commitments)))
    out.append(carr("KAT_SIGMA_COMMITMENTS", pts(keys.sigma_commitments)))
    out.append(carr("KAT_INSTANCE_COMMITMENT", pts(inst)))
    out.append(carr("KAT_PROOF", bytes(proof)))
    out.append(carr("KAT_EFWZW", pts(quad)))
    out.append(carr("KAT_PUBLIC_INPUTS", frs(public_inputs)))
    out.append(carr("KAT_PI_COMMITMENT", pts([pi_commitment])))
    out.append(carr("KAT_MULVAR_SUM", frs([plain])))
    out.append(carr("KAT_MULVAR_WEIGHTED_SUM", frs([weighted])))
    return "".join(out)


def main():
    orc.build()
    text = build()
    if "--check" in sys.argv:
        if not os.path.exists(OUT) or open(OUT).read() != text:
            sys.exit("examples/simple_example_kat.h is stale: run python tests/c/gen_simple_example_kat.py")
        print("ok")
        return
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        f.write(text)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
