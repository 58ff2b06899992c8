#!/usr/bin/env python
"""Generates halo2-aggregation_b200/csrc/field_mul_gen.cuh: the 8x32-bit Montgomery product and square of csrc/field.cuh as
straight-line carry chains (one inline-asm statement per chain, so the CC flag never crosses a statement).

  product: one level of Karatsuba over 4-limb halves (3 x 16 = 48 wide multiplications instead of 64) into a 16-limb T,
           then 8 Montgomery rows that fold m_i * p into a sliding even/odd pair of accumulators (64 wide multiplications)
  square:  a_lo^2, a_hi^2 (10 each: diagonal + doubled off-diagonal) and a_lo * a_hi (16), the same 8 rows

so a product costs 112 IMAD.WIDE and a square 100 instead of 136, paid for with IADD3 / LOP3 / SHF in issue slots the
half-rate wide multiplier leaves idle (a warp-wide IMAD.WIDE holds the pipe four cycles).

Measured on B200 (tools/micro/mulvar.cu, 4 chains per thread, 16 warps per SM): row-interleaved product 65.0 G/s (216 SASS
instructions, 136 on the multiplier), Karatsuba product 63.7 G/s (296 instructions, 126 on the multiplier: at this occupancy
the warps issue about one instruction per two cycles, so the 80 extra instructions cost what the 16 multiplications save),
square 76.6 G/s (256 instructions, 108 on the multiplier).  Hence the header carries the square only; `--with-mul` also emits
the product (kept for the record, and checked here all the same).

The SAME op lists are executed here by a small interpreter (`simulate`) against Python big integers: random and edge
operands, and every op that drops its carry asserts that there is none to drop.  `python tools/gen_field_mul.py --check`
runs that; without arguments it (re)writes the header after checking.
"""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "halo2-aggregation_b200", "csrc", "field_mul_gen.cuh")
M32 = (1 << 32) - 1

FQ = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
FR = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001


class Prog:
    """A list of statements; a statement is a carry chain (list of ops) or a plain C assignment."""

    def __init__(self):
        self.stmts = []
        self.decls = {}      # array name -> length
        self.scalars = set()

    def arr(self, name, n):
        self.decls[name] = n
        return [(name, i) for i in range(n)]

    def scalar(self, name):
        self.scalars.add(name)
        return (name, None)

    def chain(self, ops):
        self.stmts.append(("chain", ops))

    def mulwide(self, lo, hi, a, b):      # (hi:lo) = a * b          -> one IMAD.WIDE
        self.stmts.append(("mulwide", lo, hi, a, b))

    def mullo(self, d, a, b):             # d = low 32 bits of a * b -> one IMAD
        self.stmts.append(("mullo", d, a, b))

    def zero(self, d):
        self.stmts.append(("zero", d))

    def copy(self, d, s):
        self.stmts.append(("copy", d, s))

    def band(self, d, a, b):
        self.stmts.append(("and", d, a, b))

    def shl1(self, d, lo, hi):            # d = (hi:lo) << 1, upper word: funnel shift
        self.stmts.append(("shl1", d, lo, hi))

    def shr31(self, d, s):
        self.stmts.append(("shr31", d, s))


# ---------------------------------------------------------------- op semantics (the interpreter) ----------------------
def val(env, x):
    if isinstance(x, int):
        return x
    return env[x]


def simulate(prog, env):
    for st in prog.stmts:
        k = st[0]
        if k == "mulwide":
            _, lo, hi, a, b = st
            p = val(env, a) * val(env, b)
            env[lo], env[hi] = p & M32, p >> 32
        elif k == "mullo":
            _, d, a, b = st
            env[d] = (val(env, a) * val(env, b)) & M32
        elif k == "zero":
            env[st[1]] = 0
        elif k == "copy":
            env[st[1]] = val(env, st[2])
        elif k == "and":
            env[st[1]] = val(env, st[2]) & val(env, st[3])
        elif k == "shl1":
            _, d, lo, hi = st
            env[d] = ((val(env, hi) << 1) | (val(env, lo) >> 31)) & M32
        elif k == "shr31":
            env[st[1]] = val(env, st[2]) >> 31
        elif k == "chain":
            cc = None
            for op in st[1]:
                name, d, *src = op
                base = name.split(".")[0]
                uses = base in ("addc", "subc", "madc")
                sets = name.endswith(".cc")
                cin = 0
                if uses:
                    assert cc is not None, "carry used before set: %r" % (op,)
                    cin = cc
                if base in ("add", "addc"):
                    s = val(env, src[0]) + val(env, src[1]) + cin
                elif base in ("sub", "subc"):
                    s = val(env, src[0]) - val(env, src[1]) - cin
                    if sets:
                        cc = 1 if s < 0 else 0
                    elif base == "subc":
                        assert s >= 0, "borrow dropped by %r" % (op,)
                    env[d] = s & M32
                    continue
                elif base in ("mad", "madc"):
                    p = val(env, src[0]) * val(env, src[1])
                    part = (p & M32) if ".lo" in name else (p >> 32)
                    s = part + val(env, src[2]) + cin
                else:
                    raise ValueError(name)
                if sets:
                    cc = s >> 32
                else:
                    assert s >> 32 == 0, "carry dropped by %r" % (op,)
                env[d] = s & M32
        else:
            raise ValueError(k)


# ---------------------------------------------------------------- building blocks --------------------------------------
def mul4(P, out, x, y, tag):
    """out[0..7] = x[0..3] * y[0..3]; even/odd accumulators so each mad.lo/madc.hi pair is one aligned IMAD.WIDE"""
    E = P.arr("E" + tag, 8)
    O = P.arr("O" + tag, 7)     # O[k] is limb k+1
    # row 0
    P.mulwide(E[0], E[1], x[0], y[0])
    P.mulwide(E[2], E[3], x[2], y[0])
    P.mulwide(O[0], O[1], x[1], y[0])
    P.mulwide(O[2], O[3], x[3], y[0])
    # row 1: x0 y1 @1, x2 y1 @3 -> O[0..3], carry -> O[4];  x1 y1 @2, x3 y1 @4 -> E[2..5]
    P.chain([("mad.lo.cc", O[0], x[0], y[1], O[0]), ("madc.hi.cc", O[1], x[0], y[1], O[1]),
             ("madc.lo.cc", O[2], x[2], y[1], O[2]), ("madc.hi.cc", O[3], x[2], y[1], O[3]),
             ("addc", O[4], 0, 0)])
    P.chain([("mad.lo.cc", E[2], x[1], y[1], E[2]), ("madc.hi.cc", E[3], x[1], y[1], E[3]),
             ("madc.lo.cc", E[4], x[3], y[1], 0), ("madc.hi", E[5], x[3], y[1], 0)])
    # row 2: x0 y2 @2, x2 y2 @4 -> E[2..5], carry -> E[6];  x1 y2 @3, x3 y2 @5 -> O[2..5]
    P.chain([("mad.lo.cc", E[2], x[0], y[2], E[2]), ("madc.hi.cc", E[3], x[0], y[2], E[3]),
             ("madc.lo.cc", E[4], x[2], y[2], E[4]), ("madc.hi.cc", E[5], x[2], y[2], E[5]),
             ("addc", E[6], 0, 0)])
    P.chain([("mad.lo.cc", O[2], x[1], y[2], O[2]), ("madc.hi.cc", O[3], x[1], y[2], O[3]),
             ("madc.lo.cc", O[4], x[3], y[2], O[4]), ("madc.hi", O[5], x[3], y[2], 0)])
    # row 3: x0 y3 @3, x2 y3 @5 -> O[2..5], carry -> O[6];  x1 y3 @4, x3 y3 @6 -> E[4..7]
    P.chain([("mad.lo.cc", O[2], x[0], y[3], O[2]), ("madc.hi.cc", O[3], x[0], y[3], O[3]),
             ("madc.lo.cc", O[4], x[2], y[3], O[4]), ("madc.hi.cc", O[5], x[2], y[3], O[5]),
             ("addc", O[6], 0, 0)])
    P.chain([("mad.lo.cc", E[4], x[1], y[3], E[4]), ("madc.hi.cc", E[5], x[1], y[3], E[5]),
             ("madc.lo.cc", E[6], x[3], y[3], E[6]), ("madc.hi", E[7], x[3], y[3], 0)])
    # out = E + (O << 32)
    P.copy(out[0], E[0])
    P.chain([("add.cc", out[1], E[1], O[0])] + [("addc.cc", out[k], E[k], O[k - 1]) for k in range(2, 7)] +
            [("addc", out[7], E[7], O[6])])


def sqr4(P, out, x, tag):
    """out[0..7] = x[0..3]^2: 4 diagonal + 6 off-diagonal wide multiplications"""
    D = P.arr("D" + tag, 8)
    E = P.arr("E" + tag, 8)
    O = P.arr("O" + tag, 7)     # O[k] is limb k+1
    S = P.arr("S" + tag, 8)
    for i in range(4):
        P.mulwide(D[2 * i], D[2 * i + 1], x[i], x[i])
    P.mulwide(O[0], O[1], x[0], x[1])      # @1
    P.mulwide(O[2], O[3], x[0], x[3])      # @3
    P.mulwide(O[4], O[5], x[2], x[3])      # @5
    P.mulwide(E[2], E[3], x[0], x[2])      # @2
    P.mulwide(E[4], E[5], x[1], x[3])      # @4
    P.chain([("mad.lo.cc", O[2], x[1], x[2], O[2]), ("madc.hi.cc", O[3], x[1], x[2], O[3]),   # @3
             ("addc.cc", O[4], O[4], 0), ("addc", O[5], O[5], 0)])
    # S (limbs 1..6) = E + (O << 32)
    P.chain([("add.cc", S[2], E[2], O[1]), ("addc.cc", S[3], E[3], O[2]), ("addc.cc", S[4], E[4], O[3]),
             ("addc.cc", S[5], E[5], O[4]), ("addc", S[6], O[5], 0)])
    # 2S: limbs 1..7
    T2 = P.arr("W" + tag, 8)
    P.shl1(T2[1], 0, O[0])
    P.shl1(T2[2], O[0], S[2])
    for k in range(3, 7):
        P.shl1(T2[k], S[k - 1], S[k])
    P.shr31(T2[7], S[6])
    P.copy(out[0], D[0])
    P.chain([("add.cc", out[1], D[1], T2[1])] + [("addc.cc", out[k], D[k], T2[k]) for k in range(2, 7)] +
            [("addc", out[7], D[7], T2[7])])


def wide_mul(P, T, a, b):
    """T[0..15] = a[0..7] * b[0..7] by one Karatsuba level"""
    z0 = P.arr("z0", 8)
    z2 = P.arr("z2", 8)
    mul4(P, z0, a[0:4], b[0:4], "a")
    mul4(P, z2, a[4:8], b[4:8], "b")
    sa = P.arr("sa", 4)
    sb = P.arr("sb", 4)
    ca = P.scalar("ca")
    cb = P.scalar("cb")
    P.chain([("add.cc", sa[0], a[0], a[4])] + [("addc.cc", sa[k], a[k], a[k + 4]) for k in range(1, 4)] + [("addc", ca, 0, 0)])
    P.chain([("add.cc", sb[0], b[0], b[4])] + [("addc.cc", sb[k], b[k], b[k + 4]) for k in range(1, 4)] + [("addc", cb, 0, 0)])
    M = P.arr("M", 9)
    mul4(P, M, sa, sb, "c")     # writes M[0..7]
    # M += (ca ? sb : 0) << 128, (cb ? sa : 0) << 128, (ca & cb) << 256
    ma = P.scalar("ma")
    mb = P.scalar("mb")
    P.chain([("sub", ma, 0, ca)])     # 0 or 0xffffffff
    P.chain([("sub", mb, 0, cb)])
    ta = P.arr("ta", 4)
    tb = P.arr("tb", 4)
    for k in range(4):
        P.band(ta[k], sb[k], ma)
        P.band(tb[k], sa[k], mb)
    P.band(M[8], ca, cb)
    P.chain([("add.cc", M[4], M[4], ta[0])] + [("addc.cc", M[4 + k], M[4 + k], ta[k]) for k in range(1, 4)] + [("addc", M[8], M[8], 0)])
    P.chain([("add.cc", M[4], M[4], tb[0])] + [("addc.cc", M[4 + k], M[4 + k], tb[k]) for k in range(1, 4)] + [("addc", M[8], M[8], 0)])
    # z1 = M - z0 - z2   (0 <= z1 < 2^258)
    P.chain([("sub.cc", M[0], M[0], z0[0])] + [("subc.cc", M[k], M[k], z0[k]) for k in range(1, 8)] + [("subc", M[8], M[8], 0)])
    P.chain([("sub.cc", M[0], M[0], z2[0])] + [("subc.cc", M[k], M[k], z2[k]) for k in range(1, 8)] + [("subc", M[8], M[8], 0)])
    # T = z0 + (z1 << 128) + (z2 << 256)
    for k in range(4):
        P.copy(T[k], z0[k])
    P.chain([("add.cc", T[4], z0[4], M[0])] + [("addc.cc", T[4 + k], z0[4 + k], M[k]) for k in range(1, 4)] +
            [("addc.cc", T[8 + k], z2[k], M[4 + k]) for k in range(0, 5)] +
            [("addc.cc", T[13], z2[5], 0), ("addc.cc", T[14], z2[6], 0), ("addc", T[15], z2[7], 0)])


def wide_sqr(P, T, a):
    """T[0..15] = a[0..7]^2: a_lo^2 + 2 a_lo a_hi 2^128 + a_hi^2 2^256"""
    z0 = P.arr("z0", 8)
    z2 = P.arr("z2", 8)
    sqr4(P, z0, a[0:4], "a")
    sqr4(P, z2, a[4:8], "b")
    M = P.arr("M", 8)
    mul4(P, M, a[0:4], a[4:8], "c")
    M2 = P.arr("M2", 9)
    P.shl1(M2[0], 0, M[0])
    for k in range(1, 8):
        P.shl1(M2[k], M[k - 1], M[k])
    P.shr31(M2[8], M[7])
    for k in range(4):
        P.copy(T[k], z0[k])
    P.chain([("add.cc", T[4], z0[4], M2[0])] + [("addc.cc", T[4 + k], z0[4 + k], M2[k]) for k in range(1, 4)] +
            [("addc.cc", T[8 + k], z2[k], M2[4 + k]) for k in range(0, 5)] +
            [("addc.cc", T[13], z2[5], 0), ("addc.cc", T[14], z2[6], 0), ("addc", T[15], z2[7], 0)])


def mont_rows(P, R, T, p, inv):
    """R[0..7] = T / 2^256 mod p (not yet conditionally reduced: R < 2p).  Sliding pair of accumulators: X aligned at the
    current limb i, Y at limb i+1; row i adds m_i * p (even limbs of p into X, odd limbs into Y), X[0] becomes zero, the pair
    shifts by one limb (merged with the next row's odd products) and T[i+8] enters at the top."""
    A = P.arr("RA", 8)
    B = P.arr("RB", 8)
    m = P.scalar("m")
    c = P.scalar("cy")
    ones = 0xffffffff
    X, Y = A, B
    for k in range(8):
        P.copy(X[k], T[k])
    # row 0: Y = odd products alone
    P.mullo(m, X[0], inv)
    for j in range(4):
        P.mulwide(Y[2 * j], Y[2 * j + 1], p[2 * j + 1], m)
    P.chain([("mad.lo.cc", X[0], p[0], m, X[0]), ("madc.hi.cc", X[1], p[0], m, X[1]),
             ("madc.lo.cc", X[2], p[2], m, X[2]), ("madc.hi.cc", X[3], p[2], m, X[3]),
             ("madc.lo.cc", X[4], p[4], m, X[4]), ("madc.hi.cc", X[5], p[4], m, X[5]),
             ("madc.lo.cc", X[6], p[6], m, X[6]), ("madc.hi.cc", X[7], p[6], m, X[7]),
             ("addc", Y[7], Y[7], 0)])
    for i in range(1, 8):
        nx, ny = Y, X      # nx aligned at limb i (old Y), ny = old X (aligned i-1, limb 0 == 0), re-based to limb i+1
        # nx[0] += old X[1]; carry kept in c
        P.chain([("add.cc", nx[0], nx[0], ny[1]), ("addc", c, 0, 0)])
        P.mullo(m, nx[0], inv)
        # ny <- (ny >> 64) + odd products + c
        P.chain([("add.cc", c, c, ones),
                 ("madc.lo.cc", ny[0], p[1], m, ny[2]), ("madc.hi.cc", ny[1], p[1], m, ny[3]),
                 ("madc.lo.cc", ny[2], p[3], m, ny[4]), ("madc.hi.cc", ny[3], p[3], m, ny[5]),
                 ("madc.lo.cc", ny[4], p[5], m, ny[6]), ("madc.hi.cc", ny[5], p[5], m, ny[7]),
                 ("madc.lo.cc", ny[6], p[7], m, 0), ("madc.hi", ny[7], p[7], m, 0)])
        # T[i+7] enters at nx[7] (position i+7); its carry goes to ny[7] (position i+8)
        P.chain([("add.cc", nx[7], nx[7], T[i + 7]), ("addc", ny[7], ny[7], 0)])
        P.chain([("mad.lo.cc", nx[0], p[0], m, nx[0]), ("madc.hi.cc", nx[1], p[0], m, nx[1]),
                 ("madc.lo.cc", nx[2], p[2], m, nx[2]), ("madc.hi.cc", nx[3], p[2], m, nx[3]),
                 ("madc.lo.cc", nx[4], p[4], m, nx[4]), ("madc.hi.cc", nx[5], p[4], m, nx[5]),
                 ("madc.lo.cc", nx[6], p[6], m, nx[6]), ("madc.hi.cc", nx[7], p[6], m, nx[7]),
                 ("addc", ny[7], ny[7], 0)])
        X, Y = nx, ny
    # X aligned at limb 7 with X[0] == 0, Y aligned at limb 8: result (base limb 8) = X[1..7] + Y + (T[15] << 224)
    P.chain([("add.cc", R[0], X[1], Y[0])] + [("addc.cc", R[k], X[k + 1], Y[k]) for k in range(1, 7)] + [("addc", R[7], Y[7], 0)])
    P.chain([("add", R[7], R[7], T[15])])


def build(kind, p_int):
    P = Prog()
    a = P.arr("a", 8)
    b = P.arr("b", 8) if kind == "mul" else None
    T = P.arr("T", 16)
    R = P.arr("R", 8)
    p = [(p_int >> (32 * i)) & M32 for i in range(8)]
    inv = (-pow(p_int, -1, 1 << 32)) & M32
    if kind == "mul":
        wide_mul(P, T, a, b)
    else:
        wide_sqr(P, T, a)
    mont_rows(P, R, T, p, inv)
    return P


# ---------------------------------------------------------------- checking ----------------------------------------------
def limbs(x, n=8):
    return [(x >> (32 * i)) & M32 for i in range(n)]


def run(P, kind, p_int, a_int, b_int=None):
    env = {}
    for i, v in enumerate(limbs(a_int)):
        env[("a", i)] = v
    if kind == "mul":
        for i, v in enumerate(limbs(b_int)):
            env[("b", i)] = v
    simulate(P, env)
    t = sum(env[("T", i)] << (32 * i) for i in range(16))
    want_t = a_int * (b_int if kind == "mul" else a_int)
    assert t == want_t, "wide product wrong"
    r = sum(env[("R", i)] << (32 * i) for i in range(8))
    assert r < 2 * p_int, "R not below 2p"
    want = want_t * pow(1 << 256, -1, p_int) % p_int
    assert r % p_int == want, "Montgomery result wrong"


def check(n_random=3000):
    rng = random.Random(7)
    for p_int in (FQ, FR):
        edge = [0, 1, 2, p_int - 1, p_int - 2, (1 << 128) - 1, ((1 << 128) - 1) << 126 if (((1 << 128) - 1) << 126) < p_int else 5,
                (1 << 253) + (1 << 128) - 1, int("ffffffff" * 7, 16), (p_int - 1) & ~((1 << 128) - 1) | ((1 << 128) - 1),
                0x30644e72ffffffffffffffffffffffffffffffffffffffffffffffffffffffff % p_int,
                0x2fffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff,
                0x00000000ffffffff00000000ffffffff00000000ffffffff00000000ffffffff,
                0x2fffffff00000000ffffffff00000000ffffffff00000000ffffffff00000000]
        edge = [e for e in edge if e < p_int]
        for kind in ("mul", "sqr"):
            P = build(kind, p_int)
            for x in edge:
                if kind == "sqr":
                    run(P, kind, p_int, x)
                else:
                    for y in edge:
                        run(P, kind, p_int, x, y)
            for _ in range(n_random):
                x = rng.randrange(p_int)
                y = rng.randrange(p_int)
                if rng.random() < 0.3:   # limbs of all ones / zeros
                    x = sum((rng.choice([0, M32, rng.getrandbits(32)])) << (32 * i) for i in range(8)) % p_int
                    y = sum((rng.choice([0, M32, rng.getrandbits(32)])) << (32 * i) for i in range(8)) % p_int
                run(P, kind, p_int, x, y)
    return True


# ---------------------------------------------------------------- emission ----------------------------------------------
def cname(x):
    if isinstance(x, int):
        return "0x%08xu" % x
    n, i = x
    return n if i is None else "%s[%d]" % (n, i)


def emit_chain(ops):
    outs, ins = [], []     # operands: written registers ("+r" if also read before written, else "=r"), then inputs

    written, read_before = [], set()
    for op in ops:
        name, d, *src = op
        for s in src:
            if isinstance(s, int):
                continue
            if s in written:
                continue
            if s == d or True:
                pass
        # a destination read (as a source) by this or an earlier op before its first write needs "+r"
        for s in src:
            if not isinstance(s, int) and s not in written:
                read_before.add(s)
        if d not in written:
            written.append(d)
    out_ops = written
    in_ops = []
    for op in ops:
        for s in op[2:]:
            if isinstance(s, int):
                if s not in in_ops and s != 0:
                    in_ops.append(s)
            elif s not in out_ops and s not in in_ops:
                in_ops.append(s)
    index = {}
    for k, o in enumerate(out_ops):
        index[o] = k
    for k, o in enumerate(in_ops):
        index[o] = len(out_ops) + k
    lines = []
    for op in ops:
        name, d, *src = op
        args = ["%%%d" % index[d]]
        for s in src:
            if isinstance(s, int) and s == 0:
                args.append("0")
            else:
                args.append("%%%d" % index[s])
        lines.append("%s.u32 %s;" % (name, ", ".join(args)))
    # early-clobber: an output written before a later op reads an input could alias it -> mark outputs that are not
    # also inputs as "=&r" so the compiler keeps them apart from the inputs
    cons_out = []
    for o in out_ops:
        cons_out.append('"%s"(%s)' % ("+r" if o in read_before else "=&r", cname(o)))
    cons_in = ['"r"(%s)' % cname(o) for o in in_ops]
    body = '"' + '\\n\\t"\n        "'.join(lines) + '"'
    return "    asm(%s\n        : %s\n        : %s);" % (body, ", ".join(cons_out), ", ".join(cons_in) if cons_in else "")


def emit(P, fname, kind, field):
    out = []
    sig = "const uint32_t* a, const uint32_t* b" if kind == "mul" else "const uint32_t* a"
    out.append("template <> __device__ __forceinline__ void %s<%s>(uint32_t* R, %s) {" % (fname, field, sig))
    for name, n in P.decls.items():
        if name in ("a", "b", "R"):
            continue
        out.append("    uint32_t %s[%d];" % (name, n))
    for s in sorted(P.scalars):
        out.append("    uint32_t %s;" % s)
    for st in P.stmts:
        k = st[0]
        if k == "chain":
            out.append(emit_chain(st[1]))
        elif k == "mulwide":
            _, lo, hi, a, b = st
            out.append("    { const uint64_t w = (uint64_t)%s * %s; %s = (uint32_t)w; %s = (uint32_t)(w >> 32); }" %
                       (cname(a), cname(b), cname(lo), cname(hi)))
        elif k == "mullo":
            out.append("    %s = %s * %s;" % (cname(st[1]), cname(st[2]), cname(st[3])))
        elif k == "zero":
            out.append("    %s = 0;" % cname(st[1]))
        elif k == "copy":
            out.append("    %s = %s;" % (cname(st[1]), cname(st[2])))
        elif k == "and":
            out.append("    %s = %s & %s;" % (cname(st[1]), cname(st[2]), cname(st[3])))
        elif k == "shl1":
            _, d, lo, hi = st
            if isinstance(lo, int) and lo == 0:
                out.append("    %s = %s << 1;" % (cname(d), cname(hi)))
            else:
                out.append("    %s = __funnelshift_l(%s, %s, 1);" % (cname(d), cname(lo), cname(hi)))
        elif k == "shr31":
            out.append("    %s = %s >> 31;" % (cname(st[1]), cname(st[2])))
    out.append("}")
    return "\n".join(out)


HEADER = """// field_mul_gen.cuh — GENERATED by tools/gen_field_mul.py (do not edit; the generator also checks these op lists against
// Python big integers).  Montgomery product (Karatsuba over 4-limb halves: 112 IMAD.WIDE) and square (100 IMAD.WIDE) of 8 x 32-bit
// limbs for the two BN254 fields; R < 2p on return, the caller reduces once.  See csrc/field.cuh.
#pragma once
#include <cstdint>

namespace h2a {
template <int F> __device__ __forceinline__ void mont_mul_wide(uint32_t* R, const uint32_t* a, const uint32_t* b);
template <int F> __device__ __forceinline__ void mont_sqr_wide(uint32_t* R, const uint32_t* a);
"""


def main():
    check()
    if "--check" in sys.argv:
        print("ok")
        return
    parts = [HEADER + ("#define H2A_HAVE_WIDE_MUL 1\n" if "--with-mul" in sys.argv else "")]
    for field, p_int in (("0", FQ), ("1", FR)):
        if "--with-mul" in sys.argv:     # measured slower than the row-interleaved product on B200 (see the module docstring)
            parts.append(emit(build("mul", p_int), "mont_mul_wide", "mul", field))
        if "--only-mul" not in sys.argv:
            parts.append(emit(build("sqr", p_int), "mont_sqr_wide", "sqr", field))
    parts.append("}  // namespace h2a\n")
    out = OUT
    for a in sys.argv[1:]:
        if a.startswith("--out="):
            out = a[len("--out="):]
    with open(out, "w") as f:
        f.write("\n\n".join(parts))
    print("wrote", out)


if __name__ == "__main__":
    main()
