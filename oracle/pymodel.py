"""Slow Python big-int model of the BN254 ("bn256") hot-path arithmetic.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (`halo2-aggregation_b200/`)
may import this module; only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may touch `oracle/`.

PARITY UNPINNED: the reference (/root/reference) vendors none of the arithmetic
(halo2 `kzg-agg2` and halo2wrong `agg2` are un-vendored git branches,
Cargo.toml:10,12) and holds no golden vectors (src/lib.rs:43-44 is an empty
test module).  This model is pinned only to mathematical known answers
(curve order, k*G multiples, roots of unity) and to
Python's own `hashlib.blake2b`.

It restates, with plain Python integers:
  * Fq / Fr (App. A constants), Montgomery encode/decode (R = 2^256)
  * G1 affine add/double/scalar-mul on y^2 = x^3 + 3
  * MSM as a naive sum  (value of halo2 `best_multiexp`; call sites
    examples/simple-example.rs:638-640)
  * NTT as halo2 `best_fft` defines it: a[i] <- sum_j a[j] * omega^(i*j)
  * the Blake2b transcript (src/transcript.rs:58,72,105-107,122-124)
  * the GWC accumulation closed form of src/multiopen.rs:271-509
"""
import hashlib

P = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
R = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
MONT = 1 << 256
FR_S = 28
FR_GEN = 7
FR_ROOT = pow(FR_GEN, (R - 1) >> FR_S, R)
G1 = (1, 2)
B = 3


# ---------------------------------------------------------------- encodings
def to_mont(v, mod):
    return (v * MONT) % mod


def from_mont(v, mod):
    return (v * pow(MONT, -1, mod)) % mod


def le32(v):
    return int(v).to_bytes(32, "little")


def from_le(b):
    return int.from_bytes(bytes(b), "little")


def fq_mont_bytes(v):
    return le32(to_mont(v % P, P))


def fr_mont_bytes(v):
    return le32(to_mont(v % R, R))


def fq_from_mont_bytes(b):
    return from_mont(from_le(b), P)


def fr_from_mont_bytes(b):
    return from_mont(from_le(b), R)


def affine_bytes(pt):
    """x||y Montgomery LE; identity = 64 zero bytes (include/h2agg.h layout)."""
    if pt is None:
        return bytes(64)
    return fq_mont_bytes(pt[0]) + fq_mont_bytes(pt[1])


def affine_from_bytes(b):
    b = bytes(b)
    if b == bytes(64):
        return None
    return (fq_from_mont_bytes(b[:32]), fq_from_mont_bytes(b[32:64]))


# ---------------------------------------------------------------- curve
def on_curve(pt):
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - B) % P == 0


def g1_neg(pt):
    if pt is None:
        return None
    return (pt[0], (-pt[1]) % P)


def g1_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return None
        lam = (3 * x1 * x1) * pow(2 * y1, -1, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    y3 = (lam * (x1 - x3) - y1) % P
    return (x3, y3)


def g1_mul(pt, k):
    k %= R
    acc = None
    while k:
        if k & 1:
            acc = g1_add(acc, pt)
        pt = g1_add(pt, pt)
        k >>= 1
    return acc


def msm(scalars, points):
    acc = None
    for s, p in zip(scalars, points):
        acc = g1_add(acc, g1_mul(p, s))
    return acc


# ---------------------------------------------------------------- NTT
def omega_for(k):
    return pow(FR_ROOT, 1 << (FR_S - k), R)


def ntt(a, omega):
    """best_fft semantics: natural order in and out."""
    n = len(a)
    return [sum(a[j] * pow(omega, i * j, R) for j in range(n)) % R for i in range(n)]


def intt(a, omega):
    n = len(a)
    ninv = pow(n, -1, R)
    return [v * ninv % R for v in ntt(a, pow(omega, -1, R))]


def poly_eval(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % R
    return acc


# ---------------------------------------------------------------- transcript
class Blake2bTranscript:
    """halo2 Blake2bWrite/Blake2bRead + Challenge255 (SURVEY App. A encodings;
    used natively at src/transcript.rs:58,72,105-107,122-124)."""

    def __init__(self):
        self.h = hashlib.blake2b(digest_size=64, person=b"Halo2-Transcript")

    def common_point(self, pt):
        assert pt is not None
        self.h.update(b"\x01" + le32(pt[0]) + le32(pt[1]))

    def common_scalar(self, s):
        self.h.update(b"\x02" + le32(s % R))

    def squeeze_challenge(self):
        self.h.update(b"\x00")
        wide = self.h.copy().digest()
        return int.from_bytes(wide, "little") % R


def compress_point(pt):
    if pt is None:
        return bytes(32)
    b = bytearray(le32(pt[0]))
    b[31] |= (pt[1] & 1) << 7
    return bytes(b)


def sqrt_fq(a):
    y = pow(a, (P + 1) // 4, P)
    return y if y * y % P == a % P else None


def decompress_point(b):
    b = bytearray(b)
    if bytes(b) == bytes(32):
        return None
    sign = b[31] >> 7
    b[31] &= 0x7F
    x = from_le(b)
    y = sqrt_fq((x * x * x + B) % P)
    assert y is not None
    if (y & 1) != sign:
        y = P - y
    return (x, y)


# ---------------------------------------------------------------- GWC accumulation
def gwc_accumulate(queries, ws, x, u, v, omega, g1=G1):
    """Closed form of MultiopenChip::calc_witness (src/multiopen.rs:271-509).

    queries: list of (commitment_point, rotation:int, eval:int) in insertion
    order; ws: the S witness points W_i read from the proof in ascending
    rotation order.  Returns (e, f, w, zw) affine points."""
    sets = {}
    for q in queries:                       # src/multiopen.rs:19-45
        sets.setdefault(q[1], []).append(q)
    rots = sorted(sets)
    assert len(rots) == len(ws)
    omega_inv = pow(omega, -1, R)
    w_acc = zw_acc = f_acc = None
    e_acc = 0
    for rot, wi in zip(rots, ws):
        om = pow(omega, rot, R) if rot >= 0 else pow(omega_inv, -rot, R)   # :348-359
        z = x * om % R                                                     # :385-390
        zwi = g1_mul(wi, z)                                                # :393
        fb, eb = None, 0
        for (c, _r, ev) in sets[rot]:                                      # :413-462 Horner in v
            fb = g1_add(g1_mul(fb, v), c)
            eb = (eb * v + ev) % R
        w_acc = g1_add(g1_mul(w_acc, u), wi)                               # :472-476
        zw_acc = g1_add(g1_mul(zw_acc, u), zwi)                            # :478-482
        f_acc = g1_add(g1_mul(f_acc, u), fb)                               # :484-488
        e_acc = (e_acc * u + eb) % R                                       # :406-409,467-469
    e = g1_mul(g1, (-e_acc) % R)                                           # :490-492
    return e, f_acc, w_acc, zw_acc


# ---------------------------------------------------------------- setup randomness
class XorShiftRng:
    """rand_xorshift 0.3 `XorShiftRng::from_seed([u8; 16])` (Cargo.toml:16; seeded at
    examples/simple-example.rs:584-587): four little-endian u32 words x, y, z, w; all-zero seeds are replaced."""

    def __init__(self, seed16):
        seed16 = bytes(seed16)
        assert len(seed16) == 16
        self.x, self.y, self.z, self.w = (int.from_bytes(seed16[4 * i:4 * i + 4], "little") for i in range(4))
        if (self.x | self.y | self.z | self.w) == 0:
            self.x, self.y, self.z, self.w = 0x0BAD5EED, 0x0BAD5EED, 0x0BAD5EED, 0x0BAD5EED

    def next_u32(self):
        t = (self.x ^ (self.x << 11)) & 0xFFFFFFFF
        self.x, self.y, self.z = self.y, self.z, self.w
        self.w = (self.w ^ (self.w >> 19) ^ (t ^ (t >> 8))) & 0xFFFFFFFF
        return self.w

    def fill_bytes(self, n):
        out = bytearray()
        while len(out) < n:
            out += self.next_u32().to_bytes(4, "little")
        return bytes(out[:n])


REFERENCE_SETUP_SEED = bytes([0x59, 0x62, 0xbe, 0x5d, 0x76, 0x3d, 0x31, 0x8d, 0x17, 0xdb, 0x37, 0x32, 0x54, 0x06, 0xbc, 0xe5])


def setup_secret_from_seed(seed16=REFERENCE_SETUP_SEED):
    """[UPSTREAM-INFERRED] `Fr::random(rng)` of the dependency = from_bytes_wide over 64 bytes of the stream."""
    return int.from_bytes(XorShiftRng(seed16).fill_bytes(64), "little") % R


def vk_hash_from_pinned(pinned_debug):
    """The scalar `VerifierChip` absorbs for the verifying key (src/verifier.rs:341-358): Blake2b-512, personal
    "Halo2-Verify-Key", over len_le64 || bytes of `format!("{:?}", vk.pinned())`, then Fr::from_bytes_wide."""
    import hashlib
    data = pinned_debug.encode() if isinstance(pinned_debug, str) else bytes(pinned_debug)
    h = hashlib.blake2b(len(data).to_bytes(8, "little") + data, digest_size=64, person=b"Halo2-Verify-Key").digest()
    return int.from_bytes(h, "little") % R
