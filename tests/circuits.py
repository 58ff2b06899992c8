"""Workload definitions for the PLONK pipeline tests (test helpers, not product code).

`my_circuit` models the reference's only workload, `MyCircuit` of examples/simple-example.rs:316-391
(mul gate :129-142, u8 lookup :117-123, equality on instance/constant/advice :110-114, layout of
synthesize :349-390): c = constant * (a*b)^2 exposed as public input."""
import random

from oracle import plonk as pk
from oracle import pymodel as pm

R = pm.R
A, F, I = pk.ADVICE, pk.FIXED, pk.INSTANCE


def my_circuit(k=9, table_bits=8, a=3, b=5, constant=7, seed=1):
    n = 1 << k
    bf = 5
    assert n > (1 << table_bits) + bf + 1 and n >= 16
    # fixed columns: 0 constant, 1 s_mul, 2 s_lookup, 3 u8 table
    shape = pk.Shape(
        k=k, blinding_factors=bf, degree=5, num_instance=1, num_advice=2, num_fixed=4,
        advice_queries=[(0, 0), (1, 0), (0, 1)],
        fixed_queries=[(2, 0), (3, 0), (1, 0), (0, 0)],
        instance_queries=[(0, 0)],
        gates=[[(pk.OP_FIXED, 2), (pk.OP_ADVICE, 0), (pk.OP_ADVICE, 1), (pk.OP_MUL, 0), (pk.OP_ADVICE, 2), (pk.OP_NEG, 0),
                (pk.OP_ADD, 0), (pk.OP_MUL, 0)]],
        constants=[],
        lookups=[([[(pk.OP_FIXED, 0), (pk.OP_ADVICE, 0), (pk.OP_MUL, 0)]], [[(pk.OP_FIXED, 1)]])],
        perm_columns=[(I, 0, 0), (F, 0, 3), (A, 0, 0), (A, 1, 1)],
    )
    ab = a * b % R
    absq = ab * ab % R
    c = constant * absq % R
    a0, a1 = [0] * n, [0] * n
    f_const, s_mul, s_lookup, table = [0] * n, [0] * n, [0] * n, [0] * n
    for i in range(1 << table_bits):
        table[i] = i
    a0[0] = a; s_lookup[0] = 1
    a0[1] = b; s_lookup[1] = 1
    a0[2] = constant; f_const[0] = constant
    s_mul[3] = 1; a0[3] = a; a1[3] = b; a0[4] = ab
    s_mul[5] = 1; a0[5] = ab; a1[5] = ab; a0[6] = absq
    s_mul[7] = 1; a0[7] = constant; a1[7] = absq; a0[8] = c
    inst = [0] * n
    inst[0] = c
    rng = random.Random(seed)
    for col in (a0, a1):
        for r in range(n - bf, n):
            col[r] = rng.randrange(R)
    # copy constraints as cycles over (perm column position, row): 0 instance, 1 constant, 2 a0, 3 a1
    cycles = [
        [(2, 0), (2, 3)],                 # a
        [(2, 1), (3, 3)],                 # b
        [(2, 2), (1, 0), (2, 7)],         # constant
        [(2, 4), (2, 5), (3, 5)],         # ab
        [(2, 6), (3, 7)],                 # absq
        [(2, 8), (0, 0)],                 # c exposed
    ]
    return dict(shape=shape, fixed=[f_const, s_mul, s_lookup, table], cycles=cycles, instance=[inst], advice=[a0, a1],
                public_inputs=[c])


def wide_circuit(k=6, seed=2):
    """Three advice columns, an add gate with a constant and a scaled term, two lookups (one with two
    compressed inputs), seven permutation columns (three grand-product chunks, so z_last evaluations appear)."""
    n = 1 << k
    bf = 6
    rng = random.Random(seed)
    tbits = 3
    # fixed: 0 q_add, 1 table0, 2 table1a, 3 table1b, 4 q_lk, 5 konst
    shape = pk.Shape(
        k=k, blinding_factors=bf, degree=5, num_instance=1, num_advice=3, num_fixed=6,
        advice_queries=[(0, 0), (1, 0), (2, 0), (2, -1), (0, 1)],
        fixed_queries=[(0, 0), (1, 0), (2, 0), (3, 0), (4, 0), (5, 0)],
        instance_queries=[(0, 0)],
        gates=[
            # q_add * (a0 + 3*a1 + konst - a2)
            [(pk.OP_FIXED, 0), (pk.OP_ADVICE, 0), (pk.OP_ADVICE, 1), (pk.OP_SCALE, 0), (pk.OP_ADD, 0), (pk.OP_FIXED, 5), (pk.OP_ADD, 0),
             (pk.OP_ADVICE, 2), (pk.OP_NEG, 0), (pk.OP_ADD, 0), (pk.OP_MUL, 0)],
            # q_add * (a2(prev) * a0(next) - a2(prev) * a0(next))  -- exercises rotations -1 and +1, identically zero
            [(pk.OP_FIXED, 0), (pk.OP_ADVICE, 3), (pk.OP_ADVICE, 4), (pk.OP_MUL, 0), (pk.OP_ADVICE, 4), (pk.OP_ADVICE, 3), (pk.OP_MUL, 0),
             (pk.OP_NEG, 0), (pk.OP_ADD, 0), (pk.OP_MUL, 0), (pk.OP_CONST, 1), (pk.OP_MUL, 0)],
        ],
        constants=[3, 11],
        lookups=[
            ([[(pk.OP_FIXED, 4), (pk.OP_ADVICE, 0), (pk.OP_MUL, 0)]], [[(pk.OP_FIXED, 1)]]),
            ([[(pk.OP_FIXED, 4), (pk.OP_ADVICE, 0), (pk.OP_MUL, 0)], [(pk.OP_FIXED, 4), (pk.OP_ADVICE, 1), (pk.OP_MUL, 0)]],
             [[(pk.OP_FIXED, 2)], [(pk.OP_FIXED, 3)]]),
        ],
        perm_columns=[(A, 0, 0), (A, 1, 1), (A, 2, 2), (I, 0, 0), (F, 5, 5), (F, 1, 1), (F, 0, 0)],
    )
    u = shape.usable
    q_add, t0, t1a, t1b, q_lk, konst = ([0] * n for _ in range(6))
    for i in range(1 << tbits):
        t0[i] = i
        t1a[i] = i
        t1b[i] = (i * i) % (1 << tbits)
    a0, a1, a2 = [0] * n, [0] * n, [0] * n
    rows = min(u, 20)
    for r in range(rows):
        x = rng.randrange(1 << tbits)
        a0[r], a1[r] = x, (x * x) % (1 << tbits)
        konst[r] = rng.randrange(R)
        a2[r] = (a0[r] + 3 * a1[r] + konst[r]) % R
        q_add[r] = 1
        q_lk[r] = 1
    inst = [0] * n
    inst[0] = a2[0]
    # make a few cells equal so that non-trivial cycles are valid
    a0[2], a1[2] = a0[1], a1[1]
    konst[2] = konst[1]
    a2[2] = a2[1]
    cycles = [[(0, 1), (0, 2)], [(1, 1), (1, 2)], [(2, 1), (2, 2)], [(4, 1), (4, 2)], [(2, 0), (3, 0)]]
    for col in (a0, a1, a2):
        for r in range(n - bf, n):
            col[r] = rng.randrange(R)
    return dict(shape=shape, fixed=[q_add, t0, t1a, t1b, q_lk, konst], cycles=cycles, instance=[inst], advice=[a0, a1, a2],
                public_inputs=[inst[0]])


def many_rotations_circuit(k=6, seed=3):
    """`wide_circuit` with advice queries at eleven more rotations: twelve distinct evaluation points (0, +-1, ..., the
    z_last rotation), more than the eight a fixed block of point slots in front of the evaluation results could hold."""
    c = wide_circuit(k=k, seed=seed)
    sh = c["shape"]
    extra = [(0, 2), (0, -2), (1, 3), (1, -3), (2, 4), (2, -4), (0, 5), (1, -5), (2, 6), (0, 7), (1, -8)]
    c["shape"] = pk.Shape(k=sh.k, blinding_factors=sh.bf, degree=sh.degree, num_instance=sh.num_instance, num_advice=sh.num_advice,
                          num_fixed=sh.num_fixed, advice_queries=sh.advice_queries + extra, fixed_queries=sh.fixed_queries,
                          instance_queries=sh.instance_queries, gates=sh.gates, constants=sh.constants, lookups=sh.lookups,
                          perm_columns=sh.perm_columns, coset_shift=sh.coset_shift)
    return c


def setup(orc, circuit, s=0x1234567890abcdef1234567890abcdef, vk_hash=0xC0FFEE):
    shape = circuit["shape"]
    params = pk.Params(orc, shape.k, s)
    sigmas = pk.build_sigmas(shape, circuit["cycles"])
    keys = pk.Keys(orc, params, shape, circuit["fixed"], sigmas, vk_hash)
    return params, keys
