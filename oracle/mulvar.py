"""ORACLE (test infrastructure, not product code): big-integer restatement of the witness of a non-native `mul_var`,
the cells the aggregation circuit fills for every point-by-scalar multiplication of the in-circuit verifier
(src/multiopen.rs:393,443,474,480,486,492; src/vanishing.rs:181-187).

The chip that defines those cells (halo2wrong, Cargo.toml:10) is NOT in the reference tree, so this file restates its
published algorithm — Fq values as 4 limbs of 68 bits (examples/simple-example.rs:396-397,535-548), every Fq product
a*b = q*p + r witnessed through limb products against the negative wrong modulus p' = 2^272 - p, incomplete affine addition
started from an auxiliary point — in the cell order csrc/mulvar.cu documents.  PARITY UNPINNED at that boundary; the values
themselves (limbs, quotients, remainders, intermediate points, the product s*P) are plain integer facts checked here by
their defining identities (`check_record`) and against the curve arithmetic of oracle/pymodel.py."""
from . import pymodel as pm

P = pm.P
B = 68
MASK = (1 << B) - 1
NEG_P = (1 << (4 * B)) - P
BITS = 254
REC = 22
STEP = 7 * REC + 16
FINAL = 3 * REC + 8
LEN = BITS + BITS * STEP + FINAL


def limbs(v):
    return [(v >> (B * i)) & MASK for i in range(4)]


def record(a, b):
    """cells of one non-native product: a[4] b[4] q[4] r[4] t[4] v[2]; returns (cells, r)"""
    q, r = divmod(a * b, P)
    al, bl, ql, rl, pl = limbs(a), limbs(b), limbs(q), limbs(r), limbs(NEG_P)
    t = [sum(al[i] * bl[k - i] + ql[i] * pl[k - i] for i in range(k + 1)) for k in range(4)]
    u0 = t[0] + (t[1] << B) - rl[0] - (rl[1] << B)
    assert u0 >= 0 and u0 % (1 << (2 * B)) == 0
    v0 = u0 >> (2 * B)
    u1 = t[2] + (t[3] << B) - rl[2] - (rl[3] << B) + v0
    assert u1 >= 0 and u1 % (1 << (2 * B)) == 0
    v1 = u1 >> (2 * B)
    return al + bl + ql + rl + t + [v0, v1], r


def check_record(cells):
    """the identities the circuit enforces on one record (what makes the cells a valid witness)"""
    al, bl, ql, rl, t, v = cells[0:4], cells[4:8], cells[8:12], cells[12:16], cells[16:20], cells[20:22]
    pl = limbs(NEG_P)
    join = lambda ls: sum(x << (B * i) for i, x in enumerate(ls))
    a, b, q, r = join(al), join(bl), join(ql), join(rl)
    assert all(0 <= x <= MASK for x in al + bl + ql + rl)
    assert a * b == q * P + r and r < P
    for k in range(4):
        assert t[k] == sum(al[i] * bl[k - i] + ql[i] * pl[k - i] for i in range(k + 1))
    assert t[0] + (t[1] << B) - rl[0] - (rl[1] << B) == v[0] << (2 * B)
    assert t[2] + (t[3] << B) - rl[2] - (rl[3] << B) + v[0] == v[1] << (2 * B)
    # native-field relation (mod r), the other half of the CRT argument
    R = pm.R
    assert (a % R) * (b % R) % R == ((q % R) * (P % R) + r) % R
    return a, b, r


def add_cells(p1, p2):
    """T = p1 + p2 with the incomplete formula: 3 records; None when the x coordinates are equal"""
    (x1, y1), (x2, y2) = p1, p2
    dx = (x2 - x1) % P
    if dx == 0:
        return None, None
    lam = (y2 - y1) * pow(dx, -1, P) % P
    c1, _ = record(lam, dx)
    c2, l2 = record(lam, lam)
    x3 = (l2 - x1 - x2) % P
    c3, m = record(lam, (x1 - x3) % P)
    return c1 + c2 + c3, (x3, (m - y1) % P)


def double_cells(p):
    x, y = p
    c1, xx = record(x, x)
    y2 = 2 * y % P
    lam = 3 * xx * pow(y2, -1, P) % P
    c2, _ = record(lam, y2)
    c3, l2 = record(lam, lam)
    x3 = (l2 - 2 * x) % P
    c4, m = record(lam, (x - x3) % P)
    return c1 + c2 + c3 + c4, (x3, (m - y) % P)


def mulvar_witness(point, scalar, aux):
    """(result point or None, cells[LEN] as integers, status) exactly as the kernel writes them (cells past a failure are None)"""
    cells = [None] * LEN
    for i in range(BITS):
        cells[i] = (scalar >> i) & 1
    if point is None:
        return None, cells, 0xffffffff
    acc = aux
    for step in range(BITS):
        base = BITS + step * STEP
        bit = (scalar >> (BITS - 1 - step)) & 1
        cd, D = double_cells(acc)
        ca, T = add_cells(D, point)
        cells[base:base + 4 * REC] = cd
        if ca is None:
            return None, cells, 1 + step
        cells[base + 4 * REC:base + 7 * REC] = ca
        cells[base + 7 * REC:base + 7 * REC + 8] = limbs(D[0]) + limbs(D[1])
        acc = T if bit else D
        cells[base + 7 * REC + 8:base + 7 * REC + 16] = limbs(acc[0]) + limbs(acc[1])
    corr = pm.g1_neg(pm.g1_mul(aux, 1 << BITS))
    base = BITS + BITS * STEP
    cf, Q = add_cells(acc, corr)
    if cf is None:
        return None, cells, 1 + BITS
    cells[base:base + 3 * REC] = cf
    cells[base + 3 * REC:base + 3 * REC + 8] = limbs(Q[0]) + limbs(Q[1])
    return Q, cells, 0
