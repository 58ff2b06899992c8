// mulvar.cu — witness generation for the non-native `mul_var` of the aggregation circuit (row f4 of SURVEY §8), batched.
//
// The in-circuit verifier multiplies G1 points by transcript scalars with `ecc_chip.mul_var(region, point, scalar, offset)`:
// once per opening point z_i W_i (src/multiopen.rs:393), once per query in the Horner chain of commitments (:443), three
// times per rotation set for W, ZW, F (:474, :480, :486), once for E (:492) and once per quotient piece
// (src/vanishing.rs:181-187) — about 37 per aggregated proof.  The chip behind it (halo2wrong, Cargo.toml:10, not in the tree)
// represents an Fq coordinate as 4 limbs of 68 bits held in Fr cells (examples/simple-example.rs:396-397; the same packing as
// the public inputs, :535-548) and constrains every Fq product a*b = q*p + r through limb products.  Filling those cells is
// sequential big-integer code on the CPU in the reference and the Amdahl term of proving the aggregation circuit; here
// a launch pair covers every (proof, mul_var) pair of a batch: one thread per mul_var walks the ladder and shares its
// inversions, one thread per (mul_var, step) writes the records.
//
// PARITY UNPINNED: halo2wrong's exact cell layout is not visible from the reference, so the layout below is this library's
// statement of that published algorithm (integer chip with a negative wrong modulus, incomplete affine addition with an
// auxiliary point), mirrored value for value by the big-integer oracle oracle/mulvar.py.
//
// One mul_var, Q = s * P, values are Fr elements (32 bytes, Montgomery form) in this order (LEN = 43508):
//   bits[254]                      s = sum bits[i] 2^i
//   254 steps, bit 253 first; acc starts at AUX (any point; keeps the incomplete formulas away from the identity):
//     D = 2 acc:    records  x*x = xx;  lam*(2y) = 3xx;  lam*lam;  lam*(x - xD) = yD + y
//     T = D + P:    records  lam*(xP - xD) = yP - yD;  lam*lam;  lam*(xD - xT) = yT + yD
//     limbs of D (x: 4, y: 4), limbs of the new acc = bit ? T : D (8)
//   final  Q = acc + C, C = -(2^254 AUX): 3 records as for T, limbs of Q (8)
//   record (a*b = q*p + r over the integers, a, b, r < p):  a[4] b[4] q[4] r[4] t[4] v[2]  with 68-bit limbs, p' = 2^272 - p,
//     t_k = sum_{i+j=k} a_i b_j + q_i p'_j,   v_0 = (t_0 + 2^68 t_1 - r_0 - 2^68 r_1) / 2^136,
//     v_1 = (t_2 + 2^68 t_3 - r_2 - 2^68 r_3 + v_0) / 2^136   (both exact: a*b + q*p' - r = 0 mod 2^272)
// A step whose addition meets equal x coordinates (s = 0, P = +-AUX multiples, ...) cannot be witnessed with the incomplete
// formulas — the circuit would be unsatisfiable — and is reported per entry.
#include <algorithm>
#include <cstring>
#include <vector>

#include "ctx.hpp"
#include "curve.cuh"
#include "host_bn254.hpp"

using namespace h2a;

namespace {

constexpr uint32_t MV_BITS = 254, MV_REC = 22, MV_STEP = 7 * MV_REC + 16, MV_FINAL = 3 * MV_REC + 8;
constexpr uint32_t MV_LEN = MV_BITS + MV_BITS * MV_STEP + MV_FINAL;   // 43508

// p' = 2^272 - p as four 68-bit limbs of three words each
__constant__ uint32_t NEG_P_LIMBS[4][3] = {{0x278302b9u, 0xc3df73e9u, 0x2u}, {0xe978e357u, 0x2687e956u, 0xau},
                                           {0x497e7ea7u, 0xd647afbau, 0xfu}, {0x18d1ece5u, 0xfffcf9bbu, 0xfu}};
// p^-1 mod 2^256
__constant__ uint32_t P_INV_256[8] = {0x1b799c77u, 0x782df87du, 0xe1359536u, 0x6121829au, 0xe7cc257fu, 0x2750342fu, 0x6e777394u, 0x0a85dd48u};

struct W8 { uint32_t w[8]; };   // up to 256 bits (t_k < 2^140, u_k < 2^210)

__device__ __forceinline__ W8 w8_zero() { W8 r; for (int i = 0; i < 8; i++) r.w[i] = 0; return r; }
__device__ __forceinline__ void w8_add(W8& a, const W8& b) {
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) { c += (uint64_t)a.w[i] + b.w[i]; a.w[i] = (uint32_t)c; c >>= 32; }
}
__device__ __forceinline__ void w8_sub(W8& a, const W8& b) {   // a >= b
    int64_t c = 0;
    for (int i = 0; i < 8; i++) { c += (int64_t)a.w[i] - b.w[i]; a.w[i] = (uint32_t)c; c >>= 32; }
}
__device__ __forceinline__ W8 w8_shl68(const W8& a) {   // a < 2^188
    W8 r = w8_zero();
    for (int i = 0; i < 6; i++) {
        const uint64_t v = (uint64_t)a.w[i] << 4;
        r.w[i + 2] |= (uint32_t)v;
        if (i + 3 < 8) r.w[i + 3] |= (uint32_t)(v >> 32);
    }
    return r;
}
__device__ __forceinline__ W8 w8_shr136(const W8& a) {
    W8 r = w8_zero();
    for (int j = 0; j < 3; j++) r.w[j] = (a.w[j + 4] >> 8) | (a.w[j + 5] << 24);
    r.w[3] = a.w[7] >> 8;
    return r;
}
// product of two 68-bit limbs (3 words each) -> 136 bits
__device__ __forceinline__ W8 limb_mul(const uint32_t* a, const uint32_t* b) {
    uint32_t out[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 3; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 3; j++) {
            c += (uint64_t)a[i] * b[j] + out[i + j];
            out[i + j] = (uint32_t)c;
            c >>= 32;
        }
        out[i + 3] = (uint32_t)c;
    }
    W8 r = w8_zero();
    for (int i = 0; i < 6; i++) r.w[i] = out[i];
    return r;
}
// limb i (68 bits, 3 words) of a 256-bit canonical integer
__device__ __forceinline__ void limb_of(const uint32_t* v, int i, uint32_t out[3]) {
    const int bit = 68 * i, w = bit >> 5, s = bit & 31;
    uint32_t x[4];
    for (int k = 0; k < 4; k++) x[k] = w + k < 8 ? v[w + k] : 0u;
    for (int k = 0; k < 3; k++) out[k] = s ? (x[k] >> s) | (x[k + 1] << (32 - s)) : x[k];
    out[2] &= 0xfu;
}
// low 256 bits of x * y
__device__ void mul_low256(const uint32_t* x, const uint32_t* y, uint32_t* out) {
    uint32_t r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
        for (int j = 0; i + j < 8; j++) {
            c += (uint64_t)x[i] * y[j] + r[i + j];
            r[i + j] = (uint32_t)c;
            c >>= 32;
        }
    }
    for (int i = 0; i < 8; i++) out[i] = r[i];
}

// A small integer (NW 32-bit words) as an Fr element in Montgomery form: v * R mod r.  A full Montgomery product by R^2 costs 136
// wide multiplications whatever the operand; with the word-serial product run over the NW words of v only — v * X / 2^(32 NW) mod r with
// the constant X = R * 2^(32 NW) mod r — it costs 17 per word, and 22 of the 26 products of a record are such conversions.  The result
// stays below 2r (checked over random and all-ones operands in Python), so one conditional subtraction finishes it.
__constant__ uint32_t TO_MONT_X[3][8] = {
    {0x15b8b9dau, 0x93e78865u, 0xb05ea154u, 0x16df2426u, 0x302ab839u, 0x1271b743u, 0xec6c226eu, 0x06bc037eu},     // NW = 1
    {0xaf8b5a7au, 0x54c81065u, 0x97aa45f6u, 0x2b82a1efu, 0x330a5a6du, 0xbe492693u, 0x8229aff4u, 0x1847e486u},     // NW = 3
    {0xebb7ae00u, 0xee881bd8u, 0x483e9406u, 0x076add9cu, 0x184bf72du, 0xe590c2b8u, 0xff54fbf4u, 0x1df79fc3u}};    // NW = 5
template <int NW>
__device__ __forceinline__ void store_small(uint8_t* dst, const uint32_t* words) {
    static_assert(NW == 1 || NW == 3 || NW == 5, "constants exist for 1, 3 and 5 words");
    constexpr int XI = NW == 1 ? 0 : NW == 3 ? 1 : 2;
    uint32_t acc[10];
#pragma unroll
    for (int j = 0; j < 10; j++) acc[j] = 0;
#pragma unroll
    for (int i = 0; i < NW; i++) {
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) { c += (uint64_t)words[i] * TO_MONT_X[XI][j] + acc[j]; acc[j] = (uint32_t)c; c >>= 32; }
        c += acc[8]; acc[8] = (uint32_t)c; acc[9] += (uint32_t)(c >> 32);
        const uint32_t m = acc[0] * FieldConst<FR>::INV;
        c = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) { c += (uint64_t)m * FieldConst<FR>::mod(j) + acc[j]; acc[j] = (uint32_t)c; c >>= 32; }
        c += acc[8]; acc[8] = (uint32_t)c; acc[9] += (uint32_t)(c >> 32);
#pragma unroll
        for (int j = 0; j < 9; j++) acc[j] = acc[j + 1];
        acc[9] = 0;
    }
    Fr v;
#pragma unroll
    for (int j = 0; j < 8; j++) v.l[j] = acc[j];
    Fr::reduce_once(v.l);
    v.store(dst);
}
__device__ void store_limbs(uint8_t* dst, const Fq& canonical) {   // 4 cells
    for (int i = 0; i < 4; i++) {
        uint32_t l[3];
        limb_of(canonical.l, i, l);
        store_small<3>(dst + 32 * i, l);
    }
}

// one record: a * b = q * p + r  (a, b in Montgomery form, below p)
__device__ __noinline__ Fq nn_record(uint8_t* dst, const Fq& a_m, const Fq& b_m) {
    const Fq r_m = a_m * b_m;
    const Fq a = a_m.from_mont(), b = b_m.from_mont(), r = r_m.from_mont();
    uint32_t lo[8], q[8], pinv[8];
    mul_low256(a.l, b.l, lo);
    {   // lo -= r (mod 2^256), q = lo * p^-1 (mod 2^256): the exact quotient (a b - r) / p, which is below p
        int64_t c = 0;
        for (int i = 0; i < 8; i++) { c += (int64_t)lo[i] - r.l[i]; lo[i] = (uint32_t)c; c >>= 32; }
        for (int i = 0; i < 8; i++) pinv[i] = P_INV_256[i];
        mul_low256(lo, pinv, q);
    }
    uint32_t al[4][3], bl[4][3], ql[4][3], rl[4][3];
    for (int i = 0; i < 4; i++) { limb_of(a.l, i, al[i]); limb_of(b.l, i, bl[i]); limb_of(q, i, ql[i]); limb_of(r.l, i, rl[i]); }
    for (int i = 0; i < 4; i++) {
        store_small<3>(dst + 32 * i, al[i]);
        store_small<3>(dst + 32 * (4 + i), bl[i]);
        store_small<3>(dst + 32 * (8 + i), ql[i]);
        store_small<3>(dst + 32 * (12 + i), rl[i]);
    }
    W8 t[4];
    for (int k = 0; k < 4; k++) {
        t[k] = w8_zero();
        for (int i = 0; i <= k; i++) {
            uint32_t np[3] = {NEG_P_LIMBS[k - i][0], NEG_P_LIMBS[k - i][1], NEG_P_LIMBS[k - i][2]};
            w8_add(t[k], limb_mul(al[i], bl[k - i]));
            w8_add(t[k], limb_mul(ql[i], np));
        }
        store_small<5>(dst + 32 * (16 + k), t[k].w);              // t_k < 2^141
    }
    W8 v = w8_zero();
    for (int half = 0; half < 2; half++) {
        W8 u = t[2 * half];
        w8_add(u, w8_shl68(t[2 * half + 1]));
        w8_add(u, v);
        W8 rr = w8_zero(), r1 = w8_zero();
        for (int j = 0; j < 3; j++) { rr.w[j] = rl[2 * half][j]; r1.w[j] = rl[2 * half + 1][j]; }
        w8_add(rr, w8_shl68(r1));
        w8_sub(u, rr);
        v = w8_shr136(u);
        store_small<3>(dst + 32 * (20 + half), v.w);              // v < 2^74
    }
    return r_m;
}

struct MvPoint { Fq x, y; };

__device__ __forceinline__ void store_point(uint8_t* dst, const MvPoint& p) {
    store_limbs(dst, p.x.from_mont());
    store_limbs(dst + 128, p.y.from_mont());
}

// T = A + B with the incomplete formula, three records; inv_dx = 1 / (xB - xA) comes from the shared inversion
__device__ void mv_add(uint8_t* dst, const MvPoint& a, const MvPoint& b, const Fq& inv_dx, MvPoint& out) {
    const Fq dx = b.x - a.x;
    const Fq lam = (b.y - a.y) * inv_dx;
    nn_record(dst, lam, dx);
    const Fq l2 = nn_record(dst + 32 * MV_REC, lam, lam);
    out.x = l2 - a.x - b.x;
    const Fq m = nn_record(dst + 64 * MV_REC, lam, a.x - out.x);
    out.y = m - a.y;
}
// D = 2 A, four records; inv_2y = 1 / (2 yA)
__device__ void mv_double(uint8_t* dst, const MvPoint& a, const Fq& inv_2y, MvPoint& out) {
    const Fq xx = nn_record(dst, a.x, a.x);
    const Fq y2 = a.y.dbl();
    const Fq lam = (xx.dbl() + xx) * inv_2y;
    nn_record(dst + 32 * MV_REC, lam, y2);
    const Fq l2 = nn_record(dst + 64 * MV_REC, lam, lam);
    out.x = l2 - a.x.dbl();
    const Fq m = nn_record(dst + 96 * MV_REC, lam, a.x - out.x);
    out.y = m - a.y;
}

// Per-entry scratch: the 2 * 254 + 1 points of the ladder (D_s, T_s for every step, then Q), 128 bytes each — XYZZ while the
// ladder runs, then affine x || y in the first half and the denominator of the formula that produced... consumed the point in the
// second — and one 32-byte prefix product per point for the two shared inversions.
constexpr uint32_t MV_PTS = 2 * MV_BITS + 1;
constexpr size_t MV_SCRATCH = (size_t)MV_PTS * (128 + 32);

// in-place: vals[k] (stride bytes apart, count of them, none zero) <- 1 / vals[k]; prefix: count x 32 bytes of scratch
__device__ void mv_batch_invert(uint8_t* vals, size_t stride, uint32_t count, uint8_t* prefix) {
    Fq run = Fq::one();
    for (uint32_t k = 0; k < count; k++) {
        run.store(prefix + 32ull * k);
        run = run * Fq::load(vals + stride * k);
    }
    Fq inv = run.inv();
    for (uint32_t k = count; k-- > 0;) {
        const Fq v = Fq::load(vals + stride * k);
        (inv * Fq::load(prefix + 32ull * k)).store(vals + stride * k);
        inv = inv * v;
    }
}

// First kernel, one thread per mul_var; status[i] = 0 ok, 1 + step when an addition met equal x coordinates, 0xffffffff for an identity
// input.  The ladder runs in XYZZ coordinates (no inversion), ONE shared inversion turns its 509 points into affine points, a second
// one inverts the 509 denominators of the affine formulas (2y of every doubling, x2 - x1 of every addition): two field inversions per
// mul_var instead of 509.  The records are written by the second kernel from what this one leaves in the scratch.
__global__ void __launch_bounds__(32) mulvar_witness_kernel(const uint8_t* __restrict__ points, const uint8_t* __restrict__ scalars, uint32_t m,
                                                            const uint8_t* __restrict__ aux, const uint8_t* __restrict__ corr,
                                                            uint8_t* __restrict__ results, uint8_t* __restrict__ witness, uint32_t* __restrict__ status,
                                                            uint8_t* __restrict__ scratch) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    uint8_t* w = witness + 32ull * MV_LEN * i;
    uint8_t* pts = scratch + MV_SCRATCH * i;
    uint8_t* prefix = pts + 128ull * MV_PTS;
    const Affine pa = Affine::load(points + 64ull * i);
    const MvPoint P{pa.x, pa.y}, A0{Fq::load(aux), Fq::load(aux + 32)}, C{Fq::load(corr), Fq::load(corr + 32)};
    uint32_t st = 0;
    if (pa.is_identity()) st = 0xffffffffu;
    const Fr s = Fr::load(scalars + 32ull * i).from_mont();
    auto bit_of = [&](uint32_t b) { return (s.l[b >> 5] >> (b & 31)) & 1u; };
    for (uint32_t b = 0; b < MV_BITS; b++) (bit_of(b) ? Fr::one() : Fr::zero()).store(w + 32ull * b);

    // ---- the ladder in XYZZ coordinates; an addition that meets equal x coordinates ends it
    if (!st) {
        XYZZ acc = XYZZ::from_affine(Affine{A0.x, A0.y});
        for (uint32_t step = 0; step < MV_BITS; step++) {
            const XYZZ D = acc.dbl();
            if ((P.x * D.zz - D.x).is_zero()) { st = 1 + step; break; }
            XYZZ T = D;
            T.add_affine(Affine{P.x, P.y}, false);
            D.store(pts + 128ull * (2 * step));
            T.store(pts + 128ull * (2 * step + 1));
            acc = bit_of(MV_BITS - 1 - step) ? T : D;
        }
        if (!st) {
            if ((C.x * acc.zz - acc.x).is_zero()) st = 1 + MV_BITS;
            else {
                acc.add_affine(Affine{C.x, C.y}, false);
                if (acc.is_identity()) st = 1 + MV_BITS;
                else acc.store(pts + 128ull * (MV_PTS - 1));
            }
        }
    }
    MvPoint Q{Fq::zero(), Fq::zero()};
    if (!st) {
        // ---- first shared inversion: 1 / zzz of every point -> affine x, y in the first 64 bytes of its slot
        mv_batch_invert(pts + 96, 128, MV_PTS, prefix);
        for (uint32_t k = 0; k < MV_PTS; k++) {
            uint8_t* q = pts + 128ull * k;
            const Fq zi = Fq::load(q + 96), zzi = (zi * Fq::load(q + 64)).sqr();
            (Fq::load(q) * zzi).store(q);
            (Fq::load(q + 32) * zi).store(q + 32);
        }
        // ---- denominators: slot 2s: 2 y of the point step s doubles; slot 2s+1: xP - xD_s; last slot: xC - x of the final acc
        MvPoint acc = A0;
        for (uint32_t step = 0; step < MV_BITS; step++) {
            uint8_t* qd = pts + 128ull * (2 * step);
            acc.y.dbl().store(qd + 64);
            (P.x - Fq::load(qd)).store(qd + 128 + 64);
            const uint8_t* sel = bit_of(MV_BITS - 1 - step) ? qd + 128 : qd;
            acc.x = Fq::load(sel); acc.y = Fq::load(sel + 32);
        }
        (C.x - acc.x).store(pts + 128ull * (MV_PTS - 1) + 64);
        // ---- second shared inversion; the records are written by mulvar_records_kernel, one thread per step
        mv_batch_invert(pts + 64, 128, MV_PTS, prefix);
        const uint8_t* ql = pts + 128ull * (MV_PTS - 1);
        Q.x = Fq::load(ql); Q.y = Fq::load(ql + 32);
    }
    Q.x.store(results + 64ull * i);
    Q.y.store(results + 64ull * i + 32);
    status[i] = st;
}

// Second kernel: one thread per (mul_var, step).  After the ladder kernel every step's inputs are in the scratch — the affine point it
// doubles (AUX for step 0, else the D or T slot of step s-1 by the scalar's bit) and its two inverted denominators — so the 255 steps
// of a multiplication (254 ladder steps + the final correction) are independent and the records, 98 % of the work, fill the machine.
__global__ void __launch_bounds__(128) mulvar_records_kernel(const uint8_t* __restrict__ points, const uint8_t* __restrict__ scalars, uint32_t m,
                                                            const uint8_t* __restrict__ aux, const uint8_t* __restrict__ corr,
                                                            uint8_t* __restrict__ witness, const uint32_t* __restrict__ status,
                                                            const uint8_t* __restrict__ scratch) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t i = t / (MV_BITS + 1), step = t % (MV_BITS + 1);
    if (i >= m || status[i]) return;
    uint8_t* w = witness + 32ull * MV_LEN * i;
    const uint8_t* pts = scratch + MV_SCRATCH * i;
    const Fr s = Fr::load(scalars + 32ull * i).from_mont();
    auto bit_of = [&](uint32_t b) { return (s.l[b >> 5] >> (b & 31)) & 1u; };
    MvPoint acc;
    if (step == 0) { acc.x = Fq::load(aux); acc.y = Fq::load(aux + 32); }
    else {
        const uint8_t* prev = pts + 128ull * (2 * (step - 1) + bit_of(MV_BITS - step));   // step s-1 consumed bit 253 - (s-1)
        acc.x = Fq::load(prev); acc.y = Fq::load(prev + 32);
    }
    if (step < MV_BITS) {
        const Affine pa = Affine::load(points + 64ull * i);
        const MvPoint P{pa.x, pa.y};
        uint8_t* base = w + 32ull * (MV_BITS + (size_t)step * MV_STEP);
        const uint8_t* qd = pts + 128ull * (2 * step);
        MvPoint D, T;
        mv_double(base, acc, Fq::load(qd + 64), D);
        mv_add(base + 32 * 4 * MV_REC, D, P, Fq::load(qd + 128 + 64), T);
        store_point(base + 32 * 7 * MV_REC, D);
        store_point(base + 32 * (7 * MV_REC + 8), bit_of(MV_BITS - 1 - step) ? T : D);
    } else {
        const MvPoint C{Fq::load(corr), Fq::load(corr + 32)};
        uint8_t* base = w + 32ull * (MV_BITS + (size_t)MV_BITS * MV_STEP);
        MvPoint Q;
        mv_add(base, acc, C, Fq::load(pts + 128ull * (MV_PTS - 1) + 64), Q);
        store_point(base + 32 * 3 * MV_REC, Q);
    }
}

}  // namespace

extern "C" {

size_t h2a_mulvar_witness_len(void) { return MV_LEN; }

int h2a_mulvar_witness_dev(h2a_ctx* ctx, const void* d_points, const void* d_scalars, size_t m, const uint8_t aux[64], void* d_results,
                           void* d_witness, uint32_t* status_out) {
    H2A_DEVICE(ctx);
    if (!ctx || !aux || (m && (!d_points || !d_scalars || !d_results || !d_witness))) return H2A_ERR_INVALID;
    if (m == 0) return H2A_OK;
    if (m > (1u << 24)) H2A_FAIL(ctx, H2A_ERR_INVALID, "mulvar_witness: more than 2^24 entries in one call");
    namespace hh = h2a_host;
    const hh::PointA a = hh::affine_load(aux);
    if (hh::is_identity(a)) H2A_FAIL(ctx, H2A_ERR_INVALID, "mulvar_witness: the auxiliary point is the identity");
    // C = -(2^254 AUX)
    hh::PointX c = hh::px_from_affine(a);
    for (uint32_t i = 0; i < MV_BITS; i++) c = hh::px_dbl(c);
    hh::PointA ca = hh::px_to_affine(c);
    ca.y = hh::neg(ca.y);
    uint8_t small[128];
    memcpy(small, aux, 64);
    hh::affine_store(small + 64, ca);
    H2A_TRY(h2a_reserve(ctx, ctx->misc, 128 + 4 * m));
    uint8_t* d_small = (uint8_t*)ctx->misc.p;
    uint32_t* d_status = (uint32_t*)(d_small + 128);
    H2A_CUDA(ctx, cudaMemcpyAsync(d_small, small, 128, cudaMemcpyHostToDevice, ctx->stream));
    // ladder scratch (81 KB per entry): at most MV_CHUNK entries per launch
    constexpr size_t MV_CHUNK = 8192;
    H2A_TRY(h2a_reserve(ctx, ctx->aff_a, MV_SCRATCH * std::min(m, MV_CHUNK)));
    for (size_t lo = 0; lo < m; lo += MV_CHUNK) {
        const size_t cnt = std::min(MV_CHUNK, m - lo);
        mulvar_witness_kernel<<<(unsigned)((cnt + 31) / 32), 32, 0, ctx->stream>>>(
            (const uint8_t*)d_points + 64 * lo, (const uint8_t*)d_scalars + 32 * lo, (uint32_t)cnt, d_small, d_small + 64, (uint8_t*)d_results + 64 * lo,
            (uint8_t*)d_witness + 32ull * MV_LEN * lo, d_status + lo, (uint8_t*)ctx->aff_a.p);
        H2A_LAUNCH_CHECK(ctx);
        const size_t threads = cnt * (MV_BITS + 1);
        mulvar_records_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, ctx->stream>>>(
            (const uint8_t*)d_points + 64 * lo, (const uint8_t*)d_scalars + 32 * lo, (uint32_t)cnt, d_small, d_small + 64,
            (uint8_t*)d_witness + 32ull * MV_LEN * lo, d_status + lo, (const uint8_t*)ctx->aff_a.p);
        H2A_LAUNCH_CHECK(ctx);
    }
    std::vector<uint32_t> st(m);
    H2A_CUDA(ctx, cudaMemcpyAsync(st.data(), d_status, 4 * m, cudaMemcpyDeviceToHost, ctx->stream));
    H2A_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    size_t bad = 0, first = 0;
    for (size_t i = 0; i < m; i++) {
        if (status_out) status_out[i] = st[i];
        if (st[i] && !bad++) first = i;
    }
    if (bad) H2A_FAIL(ctx, H2A_ERR_INVALID, "mulvar_witness: %zu of %zu entries cannot be witnessed with incomplete additions (first: entry %zu, status %u)", bad, m, first, st[first]);
    return H2A_OK;
}

int h2a_mulvar_witness(h2a_ctx* ctx, const uint8_t* points, const uint8_t* scalars, size_t m, const uint8_t aux[64], uint8_t* out_results,
                       uint8_t* out_witness, uint32_t* status_out) {
    H2A_DEVICE(ctx);
    if (!ctx || !aux || (m && (!points || !scalars || !out_results))) return H2A_ERR_INVALID;
    if (m == 0) return H2A_OK;
    struct Dev { void* p = nullptr; ~Dev() { if (p) cudaFree(p); } } dp, ds, dr, dw;
    H2A_CUDA(ctx, cudaMalloc(&dp.p, 64 * m));
    H2A_CUDA(ctx, cudaMalloc(&ds.p, 32 * m));
    H2A_CUDA(ctx, cudaMalloc(&dr.p, 64 * m));
    H2A_CUDA(ctx, cudaMalloc(&dw.p, 32ull * MV_LEN * m));
    H2A_CUDA(ctx, cudaMemcpyAsync(dp.p, points, 64 * m, cudaMemcpyHostToDevice, ctx->stream));
    H2A_CUDA(ctx, cudaMemcpyAsync(ds.p, scalars, 32 * m, cudaMemcpyHostToDevice, ctx->stream));
    const int rc = h2a_mulvar_witness_dev(ctx, dp.p, ds.p, m, aux, dr.p, dw.p, status_out);
    H2A_CUDA(ctx, cudaMemcpyAsync(out_results, dr.p, 64 * m, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_witness) H2A_CUDA(ctx, cudaMemcpyAsync(out_witness, dw.p, 32ull * MV_LEN * m, cudaMemcpyDeviceToHost, ctx->stream));
    H2A_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return rc;
}

}  // extern "C"
