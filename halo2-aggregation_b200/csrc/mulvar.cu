// mulvar.cu — witness generation for the non-native `mul_var` of the aggregation circuit (row f4 of SURVEY §8), batched.
//
// The in-circuit verifier multiplies G1 points by transcript scalars with `ecc_chip.mul_var(region, point, scalar, offset)`:
// once per opening point z_i W_i (src/multiopen.rs:393), once per query in the Horner chain of commitments (:443), three
// times per rotation set for W, ZW, F (:474, :480, :486), once for E (:492) and once per quotient piece
// (src/vanishing.rs:181-187) — about 37 per aggregated proof.  The chip behind it (halo2wrong, Cargo.toml:10, not in the tree)
// represents an Fq coordinate as 4 limbs of 68 bits held in Fr cells (examples/simple-example.rs:396-397; the same packing as
// the public inputs, :535-548) and constrains every Fq product a*b = q*p + r through limb products.  Filling those cells is
// sequential big-integer code on the CPU in the reference and the Amdahl term of proving the aggregation circuit; here
// one thread walks one mul_var and a launch covers every (proof, mul_var) pair of a batch.
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

// an integer below 2^192 as an Fr element in Montgomery form
__device__ __forceinline__ void store_small(uint8_t* dst, const uint32_t* words, int n) {
    Fr v = Fr::zero();
    for (int i = 0; i < n; i++) v.l[i] = words[i];
    v.to_mont().store(dst);
}
__device__ void store_limbs(uint8_t* dst, const Fq& canonical) {   // 4 cells
    for (int i = 0; i < 4; i++) {
        uint32_t l[3];
        limb_of(canonical.l, i, l);
        store_small(dst + 32 * i, l, 3);
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
        store_small(dst + 32 * i, al[i], 3);
        store_small(dst + 32 * (4 + i), bl[i], 3);
        store_small(dst + 32 * (8 + i), ql[i], 3);
        store_small(dst + 32 * (12 + i), rl[i], 3);
    }
    W8 t[4];
    for (int k = 0; k < 4; k++) {
        t[k] = w8_zero();
        for (int i = 0; i <= k; i++) {
            uint32_t np[3] = {NEG_P_LIMBS[k - i][0], NEG_P_LIMBS[k - i][1], NEG_P_LIMBS[k - i][2]};
            w8_add(t[k], limb_mul(al[i], bl[k - i]));
            w8_add(t[k], limb_mul(ql[i], np));
        }
        store_small(dst + 32 * (16 + k), t[k].w, 6);
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
        store_small(dst + 32 * (20 + half), v.w, 4);
    }
    return r_m;
}

struct MvPoint { Fq x, y; };

__device__ __forceinline__ void store_point(uint8_t* dst, const MvPoint& p) {
    store_limbs(dst, p.x.from_mont());
    store_limbs(dst + 128, p.y.from_mont());
}

// T = A + B with the incomplete formula, three records; false when the x coordinates are equal
__device__ bool mv_add(uint8_t* dst, const MvPoint& a, const MvPoint& b, MvPoint& out) {
    const Fq dx = b.x - a.x;
    if (dx.is_zero()) return false;
    const Fq lam = (b.y - a.y) * dx.inv();
    nn_record(dst, lam, dx);
    const Fq l2 = nn_record(dst + 32 * MV_REC, lam, lam);
    out.x = l2 - a.x - b.x;
    const Fq m = nn_record(dst + 64 * MV_REC, lam, a.x - out.x);
    out.y = m - a.y;
    return true;
}
__device__ void mv_double(uint8_t* dst, const MvPoint& a, MvPoint& out) {
    const Fq xx = nn_record(dst, a.x, a.x);
    const Fq y2 = a.y.dbl();
    const Fq lam = (xx.dbl() + xx) * y2.inv();
    nn_record(dst + 32 * MV_REC, lam, y2);
    const Fq l2 = nn_record(dst + 64 * MV_REC, lam, lam);
    out.x = l2 - a.x.dbl();
    const Fq m = nn_record(dst + 96 * MV_REC, lam, a.x - out.x);
    out.y = m - a.y;
}

// one thread per mul_var; status[i] = 0 ok, 1 + step when an addition met equal x coordinates, 0xffffffff for an identity input
__global__ void __launch_bounds__(32) mulvar_witness_kernel(const uint8_t* __restrict__ points, const uint8_t* __restrict__ scalars, uint32_t m,
                                                            const uint8_t* __restrict__ aux, const uint8_t* __restrict__ corr,
                                                            uint8_t* __restrict__ results, uint8_t* __restrict__ witness, uint32_t* __restrict__ status) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    uint8_t* w = witness + 32ull * MV_LEN * i;
    const Affine pa = Affine::load(points + 64ull * i);
    MvPoint P{pa.x, pa.y}, acc{Fq::load(aux), Fq::load(aux + 32)}, C{Fq::load(corr), Fq::load(corr + 32)};
    uint32_t st = 0;
    if (pa.is_identity()) st = 0xffffffffu;
    const Fr s = Fr::load(scalars + 32ull * i).from_mont();
    for (uint32_t b = 0; b < MV_BITS; b++) {
        uint32_t bit[1] = {(s.l[b >> 5] >> (b & 31)) & 1u};
        store_small(w + 32ull * b, bit, 1);
    }
    for (uint32_t step = 0; step < MV_BITS && !st; step++) {
        uint8_t* base = w + 32ull * (MV_BITS + (size_t)step * MV_STEP);
        const uint32_t b = MV_BITS - 1 - step;
        MvPoint D, T;
        mv_double(base, acc, D);
        if (!mv_add(base + 32 * 4 * MV_REC, D, P, T)) { st = 1 + step; break; }
        store_point(base + 32 * 7 * MV_REC, D);
        acc = ((s.l[b >> 5] >> (b & 31)) & 1u) ? T : D;
        store_point(base + 32 * (7 * MV_REC + 8), acc);
    }
    MvPoint Q{Fq::zero(), Fq::zero()};
    if (!st) {
        uint8_t* base = w + 32ull * (MV_BITS + (size_t)MV_BITS * MV_STEP);
        if (!mv_add(base, acc, C, Q)) st = 1 + MV_BITS;
        else store_point(base + 32 * 3 * MV_REC, Q);
    }
    if (st) { Q.x = Fq::zero(); Q.y = Fq::zero(); }
    Q.x.store(results + 64ull * i);
    Q.y.store(results + 64ull * i + 32);
    status[i] = st;
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
    mulvar_witness_kernel<<<(unsigned)((m + 31) / 32), 32, 0, ctx->stream>>>((const uint8_t*)d_points, (const uint8_t*)d_scalars, (uint32_t)m, d_small,
                                                                              d_small + 64, (uint8_t*)d_results, (uint8_t*)d_witness, d_status);
    H2A_LAUNCH_CHECK(ctx);
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
