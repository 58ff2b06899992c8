// plonk_prove.cu — the prover pipeline that `create_proof` runs for the circuits the reference proves
// (examples/simple-example.rs:606-613 sample proof, :702-709 aggregation proof), laid out on the device
// around the MSM (msm.cu) and NTT (ntt.cu) kernels.  Step order = the order in which
// `VerifierChip::_verify_proof` re-reads the proof (src/verifier.rs:341-510, :718-719, src/multiopen.rs:392):
//
//   instance/advice commitments -> theta -> lookup A', S' -> beta, gamma -> permutation Z -> lookup Z ->
//   random polynomial -> y -> quotient h(X) on the extended coset, split, committed -> x -> evaluations ->
//   v, u -> one KZG witness W_i per rotation set (ascending rotation).
//
// Everything data-parallel runs in kernels of this file (expression evaluation over rows, grand-product
// prefix scans with batch inversion, quotient evaluation, Horner evaluation, v-combination, Kate division);
// the lookup permutation (bitonic sort of 256-bit keys + run matching + compaction); the host keeps only the
// Blake2b transcript.
#include <algorithm>
#include <cstring>
#include <vector>

#include "field.cuh"
#include "host_glue.hpp"
#include "dist_layout.hpp"
#include "plonk.hpp"

int h2a_ntt_run(h2a_ctx* ctx, const uint8_t* d_src, uint32_t n_in, uint8_t* d_work, uint8_t* d_dst, uint32_t log_n,
                const uint8_t omega[32], int inverse, const uint8_t* coset_shift);
int h2a_pow_tables(h2a_ctx* ctx, const uint8_t base[32], const uint8_t c[32], uint32_t n, DevBuf& out);
int h2a_pow_vector(h2a_ctx* ctx, const uint8_t base[32], const uint8_t c[32], uint32_t n, uint8_t* d_out);

namespace hh = h2a_host;
using h2a_plonk::Shape;

constexpr int MAX_COLS = 64, MAX_Q = 96, MAX_LK = 16, MAX_PERM = 48, MAX_CHUNKS = 24, MAX_PROGS = 96;
constexpr int LOG_PW = 10;
constexpr int MAX_EVALS = 512;
constexpr int MAX_EXPR = MAX_PROGS + 2 + 2 * MAX_CHUNKS + 5 * MAX_LK;   // expressions folded with y in the quotient

// Column tables of one evaluation domain (the 2^k rows, or the 2^ext_k points of the extended coset).
struct EvalTables {
    const uint8_t* col[3][MAX_COLS];  // [advice | fixed | instance][column] -> m elements
    uint32_t q_col[3][MAX_Q];
    int32_t q_rot[3][MAX_Q];
    const uint8_t* consts;
    const uint32_t* code;  // all programs, (op, arg) pairs
    uint32_t prog_off[MAX_PROGS], prog_len[MAX_PROGS];
    uint32_t mask, step;   // m - 1, and the index distance of one rotation
};

struct QuotientArgs {
    EvalTables t;
    uint32_t n_gates;                                   // programs 0 .. n_gates-1
    uint32_t n_lookups, lk_in_first[MAX_LK], lk_in_cnt[MAX_LK], lk_tab_first[MAX_LK], lk_tab_cnt[MAX_LK];
    const uint8_t *lk_pa[MAX_LK], *lk_ps[MAX_LK], *lk_z[MAX_LK];
    uint32_t n_perm, chunk_len, n_chunks, perm_type[MAX_PERM], perm_qidx[MAX_PERM];
    const uint8_t* sigma[MAX_PERM];
    const uint8_t* pz[MAX_CHUNKS];
    alignas(16) uint8_t beta_delta[MAX_PERM][32];       // beta * delta^i
    const uint8_t *l0, *llast, *lblind;
    alignas(16) uint8_t beta[32];
    alignas(16) uint8_t gamma[32];
    alignas(16) uint8_t theta[32];
    alignas(16) uint8_t y[32];
    alignas(16) uint8_t ypow[MAX_EXPR][32];             // weight of expression k of K: y^(K-1-k)
    const uint8_t *x_lo, *x_hi;                         // X_i = g * omega_ext^i as two-level tables
    const uint8_t* vanish_inv;                          // 1 / (X_i^n - 1), period vanish_mask + 1
    uint32_t vanish_mask;
    int32_t last_rot;
    uint8_t* out;
    uint32_t row_lo, row_hi;                            // rows [row_lo, row_hi) of the extended domain this launch evaluates
};

namespace dev {
using namespace h2a;

__device__ __forceinline__ Fr ld(const uint8_t* p, uint32_t i) { return Fr::load(p + 32ull * i); }

__device__ __forceinline__ Fr query(const EvalTables& t, int type, uint32_t q, uint32_t row) {
    uint32_t idx = (row + (uint32_t)(t.q_rot[type][q] * (int32_t)t.step)) & t.mask;
    return ld(t.col[type][t.q_col[type][q]], idx);
}

__device__ Fr eval_prog(const EvalTables& t, uint32_t prog, uint32_t row) {
    Fr st[16];
    int sp = 0;
    const uint32_t* c = t.code + 2 * t.prog_off[prog];
    for (uint32_t k = 0; k < t.prog_len[prog]; k++) {
        const uint32_t op = c[2 * k], arg = c[2 * k + 1];
        switch (op) {
            case 0: st[sp++] = ld(t.consts, arg); break;
            case 1: st[sp++] = query(t, 0, arg, row); break;
            case 2: st[sp++] = query(t, 1, arg, row); break;
            case 3: st[sp++] = query(t, 2, arg, row); break;
            case 4: st[sp - 1] = st[sp - 1].neg(); break;
            case 5: st[sp - 2] = st[sp - 2] + st[sp - 1]; sp--; break;
            case 6: st[sp - 2] = st[sp - 2] * st[sp - 1]; sp--; break;
            default: st[sp - 1] = st[sp - 1] * ld(t.consts, arg); break;
        }
    }
    return st[0];
}

__device__ Fr compress(const EvalTables& t, uint32_t first, uint32_t cnt, const Fr& theta, uint32_t row) {
    if (cnt == 0) return Fr::zero();
    Fr acc = eval_prog(t, first, row);
    for (uint32_t p = 1; p < cnt; p++) acc = acc * theta + eval_prog(t, first + p, row);
    return acc;
}

// out[row] = theta-compression of programs [first, first+cnt) at every row     (src/lookup.rs:229-262)
__global__ void __launch_bounds__(128) compress_kernel(const EvalTables* t, uint32_t first, uint32_t cnt, const uint8_t* theta,
                                                       uint32_t m, uint8_t* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    compress(*t, first, cnt, Fr::load(theta), i).store(out + 32ull * i);
}

// h(X_i) numerator folded with y, divided by the vanishing polynomial, for every point of the extended coset.
// Expression order: gates, permutation 1-4 (src/permutation.rs:211-321), five per lookup (src/lookup.rs:190-310).
// The fold sum_k e_k y^(K-1-k) is taken with the weights w_k = y^(K-1-k) (a.ypow, from the host) instead of a Horner
// chain, so that the expressions sharing the factor l_0, l_last or 1 - (l_last + l_blind) are summed first and meet
// that factor once: about 170 field products per point instead of 240 for the k=20 profile.
template <int MIN_BLOCKS>
__global__ void __launch_bounds__(128, MIN_BLOCKS) quotient_kernel(const QuotientArgs* ap) {
    const QuotientArgs& a = *ap;
    const EvalTables& t = a.t;
    const uint32_t i = a.row_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.row_hi) return;
    const Fr beta = Fr::load(a.beta), gamma = Fr::load(a.gamma), one = Fr::one();
    uint32_t k = 0;                                       // expression counter
    auto w = [&]() { return Fr::load(a.ypow[k++]); };
    Fr acc = Fr::zero(), s_l0 = Fr::zero(), s_last = Fr::zero(), s_om = Fr::zero();
    for (uint32_t g = 0; g < a.n_gates; g++) acc = acc + w() * eval_prog(t, g, i);
    const uint32_t nxt = (i + t.step) & t.mask, prv = (i - t.step) & t.mask;
    if (a.n_chunks) {
        s_l0 = w() * (one - ld(a.pz[0], i));
        Fr zl = ld(a.pz[a.n_chunks - 1], i);
        s_last = w() * (zl.sqr() - zl);
        const uint32_t lastidx = (i + (uint32_t)(a.last_rot * (int32_t)t.step)) & t.mask;
        for (uint32_t c = 1; c < a.n_chunks; c++) s_l0 = s_l0 + w() * (ld(a.pz[c], i) - ld(a.pz[c - 1], lastidx));
        Fr x = ld(a.x_lo, i & ((1u << LOG_PW) - 1u));
        if (i >> LOG_PW) x = x * ld(a.x_hi, i >> LOG_PW);
        for (uint32_t c = 0; c < a.n_chunks; c++) {
            Fr left = ld(a.pz[c], nxt), right = ld(a.pz[c], i);
            const uint32_t hi = min((c + 1) * a.chunk_len, a.n_perm);
            for (uint32_t q = c * a.chunk_len; q < hi; q++) {
                Fr val = query(t, a.perm_type[q], a.perm_qidx[q], i);
                left = left * (beta * ld(a.sigma[q], i) + val + gamma);
                right = right * (Fr::load(a.beta_delta[q]) * x + val + gamma);
            }
            s_om = s_om + w() * (left - right);
        }
    }
    if (a.n_lookups) {
        const Fr theta = Fr::load(a.theta);
        for (uint32_t l = 0; l < a.n_lookups; l++) {
            Fr z = ld(a.lk_z[l], i), zn = ld(a.lk_z[l], nxt), pa = ld(a.lk_pa[l], i), pap = ld(a.lk_pa[l], prv), ps = ld(a.lk_ps[l], i);
            s_l0 = s_l0 + w() * (one - z);
            s_last = s_last + w() * (z.sqr() - z);
            Fr ci = compress(t, a.lk_in_first[l], a.lk_in_cnt[l], theta, i);
            Fr ct = compress(t, a.lk_tab_first[l], a.lk_tab_cnt[l], theta, i);
            Fr left = (pa + beta) * (ps + gamma) * zn, right = (ci + beta) * (ct + gamma) * z;
            s_om = s_om + w() * (left - right);
            const Fr d = pa - ps;
            s_l0 = s_l0 + w() * d;
            s_om = s_om + w() * (d * (pa - pap));
        }
    }
    const Fr llast = ld(a.llast, i);
    acc = acc + ld(a.l0, i) * s_l0 + llast * s_last + (one - (llast + ld(a.lblind, i))) * s_om;
    (acc * ld(a.vanish_inv, i & a.vanish_mask)).store(a.out + 32ull * i);
}

struct PermTermArgs {
    const uint8_t* vals[8];
    const uint8_t* sigma[8];
    alignas(16) uint8_t beta_delta[8][32];
    uint32_t ncols;
};
// num[i] = prod_c (beta delta^c omega^i + gamma + v_c[i]);  den[i] = prod_c (beta sigma_c[i] + gamma + v_c[i])
__global__ void __launch_bounds__(128) perm_terms_kernel(PermTermArgs a, const uint8_t* beta_, const uint8_t* gamma_,
                                                         const uint8_t* omega_pows, uint32_t n, uint8_t* num, uint8_t* den) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Fr beta = Fr::load(beta_), gamma = Fr::load(gamma_), w = ld(omega_pows, i);
    Fr nu = Fr::one(), de = Fr::one();
    for (uint32_t c = 0; c < a.ncols; c++) {
        Fr v = ld(a.vals[c], i);
        nu = nu * (Fr::load(a.beta_delta[c]) * w + gamma + v);
        de = de * (beta * ld(a.sigma[c], i) + gamma + v);
    }
    nu.store(num + 32ull * i);
    de.store(den + 32ull * i);
}
// rows < usable: num = (A+beta)(S+gamma), den = (A'+beta)(S'+gamma); other rows 1      (src/lookup.rs:81-106)
__global__ void __launch_bounds__(128) lookup_terms_kernel(const uint8_t* A, const uint8_t* S, const uint8_t* pa, const uint8_t* ps,
                                                           const uint8_t* beta_, const uint8_t* gamma_, uint32_t usable, uint32_t n,
                                                           uint8_t* num, uint8_t* den) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr nu = Fr::one(), de = Fr::one();
    if (i < usable) {
        const Fr beta = Fr::load(beta_), gamma = Fr::load(gamma_);
        nu = (ld(A, i) + beta) * (ld(S, i) + gamma);
        de = (ld(pa, i) + beta) * (ld(ps, i) + gamma);
    }
    nu.store(num + 32ull * i);
    de.store(den + 32ull * i);
}

// In-place inversion, 16 elements per thread with one field inversion (Montgomery's trick); zeros stay zero.
constexpr int INV_CHUNK = 16;
__global__ void __launch_bounds__(128) batch_inverse_kernel(uint8_t* a, uint32_t n) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t lo = t * INV_CHUNK;
    if (lo >= n) return;
    uint32_t cnt = min((uint32_t)INV_CHUNK, n - lo);
    Fr pre[INV_CHUNK];
    Fr run = Fr::one();
    for (uint32_t k = 0; k < cnt; k++) {
        pre[k] = run;
        Fr v = ld(a, lo + k);
        if (!v.is_zero()) run = run * v;
    }
    Fr inv = run.inv();
    for (int k = (int)cnt - 1; k >= 0; k--) {
        Fr v = ld(a, lo + k);
        if (v.is_zero()) continue;
        (inv * pre[k]).store(a + 32ull * (lo + k));
        inv = inv * v;
    }
}
__global__ void mul_arrays_kernel(const uint8_t* a, const uint8_t* b, uint32_t n, uint8_t* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) (ld(a, i) * ld(b, i)).store(out + 32ull * i);
}

// ---- inclusive prefix product, in place: tiles of 2048 (256 threads x 8), recursion over the tile totals
constexpr int SCAN_T = 256, SCAN_I = 8, SCAN_TILE = SCAN_T * SCAN_I;
__global__ void __launch_bounds__(SCAN_T) scan_mul_tiles_kernel(uint8_t* data, uint32_t n, uint8_t* totals) {
    __shared__ __align__(16) uint8_t sh[SCAN_T * 32];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_I;
    Fr v[SCAN_I];
    Fr run = Fr::one();
#pragma unroll
    for (int k = 0; k < SCAN_I; k++) {
        if (base + k < n) run = run * ld(data, base + k);
        v[k] = run;
    }
    run.store(sh + 32 * threadIdx.x);
    __syncthreads();
    for (int d = 1; d < SCAN_T; d <<= 1) {  // Hillis-Steele over the thread totals
        Fr mine = Fr::load(sh + 32 * threadIdx.x), other = Fr::one();
        const bool has = (int)threadIdx.x >= d;
        if (has) other = Fr::load(sh + 32 * (threadIdx.x - d));
        __syncthreads();
        if (has) (mine * other).store(sh + 32 * threadIdx.x);
        __syncthreads();
    }
    Fr off = threadIdx.x ? Fr::load(sh + 32 * (threadIdx.x - 1)) : Fr::one();
#pragma unroll
    for (int k = 0; k < SCAN_I; k++)
        if (base + k < n) (v[k] * off).store(data + 32ull * (base + k));
    if (threadIdx.x == SCAN_T - 1) Fr::load(sh + 32 * (SCAN_T - 1)).store(totals + 32ull * blockIdx.x);
}
__global__ void scan_mul_apply_kernel(uint8_t* data, uint32_t n, const uint8_t* totals_inclusive) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t tile = i / SCAN_TILE;
    if (i >= n || tile == 0) return;
    (ld(data, i) * ld(totals_inclusive, tile - 1)).store(data + 32ull * i);
}
// z[0] = init, z[i] = init * incl[i-1]   (grand product column from the inclusive scan of the ratios)
__global__ void shift_scale_kernel(const uint8_t* incl, const uint8_t* init_, uint32_t n, uint8_t* z) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr init = Fr::load(init_);
    (i ? init * ld(incl, i - 1) : init).store(z + 32ull * i);
}

// ---- chunked Horner: thread t folds coefficients [t*L, (t+1)*L) at the point z
constexpr int HORNER_L = 64;
__device__ __forceinline__ Fr chunk_value(const uint8_t* coef, uint32_t n, uint32_t t, const Fr& z) {
    Fr acc = Fr::zero();
    const uint32_t lo = t * HORNER_L, hi = min(lo + HORNER_L, n);
    for (uint32_t i = hi; i-- > lo;) acc = acc * z + ld(coef, i);
    return acc;
}
__global__ void __launch_bounds__(128) chunk_values_kernel(const uint8_t* coef, uint32_t n, const uint8_t* z_, uint8_t* out) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t * HORNER_L >= n) return;
    chunk_value(coef, n, t, Fr::load(z_)).store(out + 32ull * t);
}
// All evaluations of a proof in two launches: request r = (coefficients, point, result slot).
// p(z) = sum_t C_t (z^L)^t over the chunk values: per request one block, strided Horner then a shared-memory tree.
struct EvalReq {
    const uint8_t* coef;
    const uint8_t* z;
    uint8_t* out;
};
__global__ void __launch_bounds__(128) chunk_values_multi_kernel(const EvalReq* reqs, uint32_t n, uint32_t nchunks, uint8_t* chunks) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nchunks) return;
    const EvalReq r = reqs[blockIdx.y];
    chunk_value(r.coef, n, t, Fr::load(r.z)).store(chunks + 32ull * ((size_t)blockIdx.y * nchunks + t));
}
__global__ void __launch_bounds__(256) fold_chunks_multi_kernel(const EvalReq* reqs, uint32_t nchunks, const uint8_t* chunks) {
    __shared__ __align__(16) uint8_t sh[256 * 32];
    const EvalReq r = reqs[blockIdx.x];
    const uint8_t* C = chunks + 32ull * (size_t)blockIdx.x * nchunks;
    Fr z = Fr::load(r.z), zl = z;
#pragma unroll 1
    for (int k = 0; k < 6; k++) zl = zl.sqr();  // z^64
    Fr zb = zl;
#pragma unroll 1
    for (int k = 0; k < 8; k++) zb = zb.sqr();  // zl^256
    Fr acc = Fr::zero();
    uint32_t cnt = nchunks > threadIdx.x ? (nchunks - threadIdx.x + 255) / 256 : 0;
    for (uint32_t m = cnt; m-- > 0;) acc = acc * zb + ld(C, threadIdx.x + 256 * m);
    uint32_t e[1] = {threadIdx.x};
    if (cnt) acc = acc * zl.pow_limbs(e, 8);
    acc.store(sh + 32 * threadIdx.x);
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) (Fr::load(sh + 32 * threadIdx.x) + Fr::load(sh + 32 * (threadIdx.x + s))).store(sh + 32 * threadIdx.x);
        __syncthreads();
    }
    if (threadIdx.x == 0) Fr::load(sh).store(r.out);
}

// acc[i] = acc[i] * v + p[i]                       (Horner in v over the polynomials of one rotation set)
__global__ void axpy_kernel(uint8_t* acc, const uint8_t* p, const uint8_t* v_, uint32_t n, int first) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (first) ld(p, i).store(acc + 32ull * i);
    else (ld(acc, i) * Fr::load(v_) + ld(p, i)).store(acc + 32ull * i);
}

// ---- Kate division q = (p - p(z)) / (X - z): q_j = sum_{i>j} p_i z^(i-j-1)
// carries S_t = sum_{i >= (t+1)L} p_i z^(i-(t+1)L) from the chunk values: one block of 128 threads
__global__ void __launch_bounds__(128) kate_carry_kernel(const uint8_t* C, uint32_t nchunks, const uint8_t* z_, uint8_t* S) {
    __shared__ __align__(16) uint8_t shD[128 * 32];
    __shared__ __align__(16) uint8_t shT[128 * 32];
    Fr z = Fr::load(z_), zl = z;
#pragma unroll 1
    for (int k = 0; k < 6; k++) zl = zl.sqr();                    // z^L
    const uint32_t M = (nchunks + 127) / 128;                     // chunks per thread
    const uint32_t lo = threadIdx.x * M, hi = min(lo + M, nchunks);
    Fr D = Fr::zero();                                            // D_s = sum_{j<M} C_{lo+j} zl^j
    for (uint32_t t = hi; t-- > lo && t < nchunks;) D = D * zl + ld(C, t);
    D.store(shD + 32 * threadIdx.x);
    __syncthreads();
    if (threadIdx.x == 0) {                                       // T_s = carry into the top chunk of group s
        uint32_t e[1] = {M};
        Fr zlm = zl.pow_limbs(e, 32 - __clz(M | 1));
        Fr T = Fr::zero();
        for (int s = 127; s >= 0; s--) {
            T.store(shT + 32 * s);
            T = Fr::load(shD + 32 * s) + zlm * T;
        }
    }
    __syncthreads();
    if (lo < nchunks) {
        Fr carry = Fr::load(shT + 32 * threadIdx.x);
        for (uint32_t t = hi; t-- > lo;) {
            carry.store(S + 32ull * t);
            carry = ld(C, t) + zl * carry;
        }
    }
}
__global__ void __launch_bounds__(128) kate_quotient_kernel(const uint8_t* p, uint32_t n, const uint8_t* z_, const uint8_t* S, uint8_t* q) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lo = t * HORNER_L;
    if (lo >= n) return;
    const uint32_t hi = min(lo + HORNER_L, n);
    Fr z = Fr::load(z_), carry = ld(S, t);
    for (uint32_t i = hi; i-- > lo;) {
        Fr pi = ld(p, i);
        carry.store(q + 32ull * i);
        carry = pi + z * carry;
    }
}

// ---- lookup permutation on the device (halo2 `permute_expression_pair`, src/lookup.rs:49-79 reads its output)
// keys are canonical 256-bit integers so that their order is the numeric order of the field elements
__device__ __forceinline__ bool key_less(const Fr& a, const Fr& b) {
#pragma unroll
    for (int i = 7; i >= 0; i--) {
        if (a.l[i] != b.l[i]) return a.l[i] < b.l[i];
    }
    return false;
}
__global__ void lookup_keys_kernel(const uint8_t* src, uint32_t u, uint32_t n, uint8_t* keys) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr k;
    if (i < u) k = ld(src, i).from_mont();
    else {
#pragma unroll
        for (int j = 0; j < 8; j++) k.l[j] = 0xffffffffu;  // sentinel above every field element
    }
    k.store(keys + 32ull * i);
}
__global__ void bitonic_step_kernel(uint8_t* keys, uint32_t n, uint32_t j, uint32_t kk) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t p = i ^ j;
    if (i >= n || p <= i) return;
    Fr a = ld(keys, i), b = ld(keys, p);
    const bool up = (i & kk) == 0;
    if (key_less(b, a) == up) {
        b.store(keys + 32ull * i);
        a.store(keys + 32ull * p);
    }
}
// shared-memory tail of a bitonic stage: all steps with j < 1024 of stage kk in one launch (tile of 2048 keys)
__global__ void __launch_bounds__(1024) bitonic_tail_kernel(uint8_t* keys, uint32_t n, uint32_t j_start, uint32_t kk) {
    extern __shared__ uint4 sh4[];
    uint4* lo = sh4;
    uint4* hi = sh4 + 2048;
    const uint32_t base = blockIdx.x * 2048;
    for (uint32_t t = threadIdx.x; t < 2048; t += 1024) {
        const uint4* g = (const uint4*)(keys + 32ull * (base + t));
        lo[t] = g[0];
        hi[t] = g[1];
    }
    __syncthreads();
    for (uint32_t j = j_start; j > 0; j >>= 1) {
        const uint32_t t = threadIdx.x;
        const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // element with bit j clear
        const uint32_t p = i | j;
        Fr a, b;
        uint4 x = lo[i], y = hi[i];
        a.l[0] = x.x; a.l[1] = x.y; a.l[2] = x.z; a.l[3] = x.w; a.l[4] = y.x; a.l[5] = y.y; a.l[6] = y.z; a.l[7] = y.w;
        x = lo[p]; y = hi[p];
        b.l[0] = x.x; b.l[1] = x.y; b.l[2] = x.z; b.l[3] = x.w; b.l[4] = y.x; b.l[5] = y.y; b.l[6] = y.z; b.l[7] = y.w;
        const bool up = ((base + i) & kk) == 0;
        if (key_less(b, a) == up) {
            lo[i] = make_uint4(b.l[0], b.l[1], b.l[2], b.l[3]); hi[i] = make_uint4(b.l[4], b.l[5], b.l[6], b.l[7]);
            lo[p] = make_uint4(a.l[0], a.l[1], a.l[2], a.l[3]); hi[p] = make_uint4(a.l[4], a.l[5], a.l[6], a.l[7]);
        }
        __syncthreads();
    }
    for (uint32_t t = threadIdx.x; t < 2048; t += 1024) {
        uint4* g = (uint4*)(keys + 32ull * (base + t));
        g[0] = lo[t];
        g[1] = hi[t];
    }
}
// ---- counting sort for small keys (range tables and the like): every key < 2^COUNT_BITS fits one counter
constexpr uint32_t COUNT_BITS = 20;
// stats[0] |= OR of limbs 1..7 of all keys, stats[1] = max of limb 0   (decides which sort applies)
__global__ void key_stats_kernel(const uint8_t* keys, uint32_t u, uint32_t* stats) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t hi = 0, lo = 0;
    if (i < u) {
        Fr k = ld(keys, i);
        hi = k.l[1] | k.l[2] | k.l[3] | k.l[4] | k.l[5] | k.l[6] | k.l[7];
        lo = k.l[0];
    }
    hi = __reduce_or_sync(0xffffffffu, hi);
    lo = __reduce_max_sync(0xffffffffu, lo);
    if ((threadIdx.x & 31) == 0) {
        if (hi) atomicOr(&stats[0], hi);
        atomicMax(&stats[1], lo);
    }
}
// lookup inputs repeat (that is what a lookup is for): one atomic per distinct key of a warp, not one per lane
__global__ void count_keys_kernel(const uint8_t* keys, uint32_t u, uint32_t* hist) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31u;
    const bool live = i < u;
    const uint32_t key = live ? ld(keys, i).l[0] : (0x80000000u | lane);   // keys are < 2^COUNT_BITS
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    if (live && lane == (uint32_t)__ffs(peers) - 1u) atomicAdd(&hist[key], (uint32_t)__popc(peers));
}
// out[p] = the key whose run covers position p (offsets = exclusive prefix of the counters), sentinel beyond u
__global__ void expand_counts_kernel(const uint32_t* offsets, uint32_t bins, uint32_t u, uint32_t n, uint8_t* out) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    Fr k;
    if (p < u) {
        uint32_t lo = 0, hi = bins;  // largest v with offsets[v] <= p
        while (hi - lo > 1u) {
            uint32_t mid = (lo + hi) >> 1;
            if (offsets[mid] <= p) lo = mid; else hi = mid;
        }
        k = Fr::zero();
        k.l[0] = lo;
    } else {
#pragma unroll
        for (int j = 0; j < 8; j++) k.l[j] = 0xffffffffu;
    }
    k.store(out + 32ull * p);
}

// first[row] = 1 when row starts a run of equal inputs; used[j] = 1 for the table entry matched to that run
__global__ void lookup_match_kernel(const uint8_t* ka, const uint8_t* ks, uint32_t u, uint32_t* rep, uint32_t* unused, uint32_t* err) {
    uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= u) return;
    Fr a = ld(ka, row);
    bool first = row == 0;
    if (!first) {
        Fr prev = ld(ka, row - 1);
        first = !(prev == a);
    }
    rep[row] = first ? 0u : 1u;
    if (!first) return;
    uint32_t lo = 0, hi = u;  // lower_bound of a in ks[0..u)
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (key_less(ld(ks, mid), a)) lo = mid + 1; else hi = mid;
    }
    if (lo < u && ld(ks, lo) == a) unused[lo] = 0u;   // unused[] starts as all ones
    else atomicExch(err, 1u);
}
__global__ void fill_u32_kernel(uint32_t* a, uint32_t n, uint32_t v) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}
// exclusive scan of u32 flags, tiles of 2048: sums, single-block scan of the sums, apply
__global__ void __launch_bounds__(256) u32_tile_sums_kernel(const uint32_t* a, uint32_t n, uint32_t* sums) {
    __shared__ uint32_t sh[256];
    uint32_t base = blockIdx.x * 2048 + threadIdx.x * 8, s = 0;
    for (int k = 0; k < 8; k++) s += (base + k < n) ? a[base + k] : 0u;
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) {
        if ((int)threadIdx.x < d) sh[threadIdx.x] += sh[threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[blockIdx.x] = sh[0];
}
__global__ void u32_scan_sums_kernel(uint32_t* sums, uint32_t tiles, uint32_t* total) {  // one thread: tiles <= 2^15
    uint32_t run = 0;
    for (uint32_t t = 0; t < tiles; t++) { uint32_t v = sums[t]; sums[t] = run; run += v; }
    *total = run;
}
__global__ void __launch_bounds__(256) u32_tile_apply_kernel(const uint32_t* a, uint32_t n, const uint32_t* sums, uint32_t* out) {
    __shared__ uint32_t sh[256];
    uint32_t base = blockIdx.x * 2048 + threadIdx.x * 8, v[8], s = 0;
    for (int k = 0; k < 8; k++) { v[k] = (base + k < n) ? a[base + k] : 0u; s += v[k]; }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1) {
        uint32_t add = (int)threadIdx.x >= d ? sh[threadIdx.x - d] : 0u;
        __syncthreads();
        sh[threadIdx.x] += add;
        __syncthreads();
    }
    uint32_t ex = sh[threadIdx.x] - s + sums[blockIdx.x];
    for (int k = 0; k < 8; k++) { if (base + k < n) out[base + k] = ex; ex += v[k]; }
}
// leftover[rank] = k-th table entry that no input run consumed (ascending)
__global__ void lookup_leftover_kernel(const uint8_t* ks, const uint32_t* unused, const uint32_t* rank, uint32_t u, uint8_t* leftover) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < u && unused[j]) ld(ks, j).store(leftover + 32ull * rank[j]);
}
// A' = sorted inputs; S' = the run's own value on first rows, left-over table entries on repeated rows, handed out
// from the last repeated row backwards; both back in Montgomery form
__global__ void lookup_assign_kernel(const uint8_t* ka, const uint8_t* leftover, const uint32_t* rep, const uint32_t* rep_rank,
                                     const uint32_t* n_rep, uint32_t u, uint8_t* pa, uint8_t* ps) {
    uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= u) return;
    Fr a = ld(ka, row);
    a.to_mont().store(pa + 32ull * row);
    Fr s = a;
    if (rep[row]) s = ld(leftover, *n_rep - 1u - rep_rank[row]);
    s.to_mont().store(ps + 32ull * row);
}

// ---- KZG setup: Lagrange-basis scalars and fixed-base multiplication
// den[i] = n * (s - omega^i)
__global__ void lagrange_den_kernel(const uint8_t* omega_pows, const uint8_t* s_, const uint8_t* n_, uint32_t n, uint8_t* den) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) (Fr::load(n_) * (Fr::load(s_) - ld(omega_pows, i))).store(den + 32ull * i);
}
// out[i] = omega^i * (s^n - 1) * den_inv[i]
__global__ void lagrange_num_kernel(const uint8_t* omega_pows, const uint8_t* snm1_, const uint8_t* den_inv, uint32_t n, uint8_t* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) (ld(omega_pows, i) * Fr::load(snm1_) * ld(den_inv, i)).store(out + 32ull * i);
}

}  // namespace dev

#include "curve.cuh"
namespace dev {
// out[i] = scalars[i] * G with a table of 2^j * G (affine): mixed additions only, one inversion per point
__global__ void __launch_bounds__(128) fixed_base_mul_kernel(const uint8_t* table, const uint8_t* scalars, uint32_t n, uint8_t* out) {
    __shared__ __align__(16) uint8_t tab[254 * 64];
    for (uint32_t k = threadIdx.x; k < 254 * 4; k += blockDim.x) ((uint4*)tab)[k] = ((const uint4*)table)[k];
    __syncthreads();
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr sc = ld(scalars, i).from_mont();
    XYZZ acc = XYZZ::identity();
#pragma unroll 1
    for (int limb = 0; limb < 8; limb++) {
        uint32_t w = sc.l[limb];  // limb index is the loop counter of a fully dynamic loop: local memory by design
#pragma unroll 1
        for (int b = 0; b < 32 && 32 * limb + b < 254; b++)
            if ((w >> b) & 1) acc.add_affine(Affine::load(tab + 64 * (32 * limb + b)), false);
    }
    Affine o;
    if (acc.is_identity()) {
        o.x = Fq::zero();
        o.y = Fq::zero();
    } else {
        Fq zi = acc.zzz.inv();
        Fq zzi = (zi * acc.zz).sqr();
        o.x = acc.x * zzi;
        o.y = acc.y * zi;
    }
    o.store(out + 64ull * i);
}
}  // namespace dev

// ====================================================================== host side
struct Poly3 {
    uint8_t *lag = nullptr, *coef = nullptr, *ext = nullptr;
};

struct ProverState {
    uint8_t* arena = nullptr;
    size_t arena_bytes = 0, used = 0;
    const h2a_bases *g = nullptr, *g_lagrange = nullptr;
    uint8_t coset[32];
    std::vector<Poly3> fixed, sigma, advice, instance, pz;
    struct Lk { uint8_t *A, *S; Poly3 pa, ps, z; };
    std::vector<Lk> lk;
    uint8_t *l0 = nullptr, *llast = nullptr, *lblind = nullptr;          // extended
    uint8_t *random_coef = nullptr, *h_ext = nullptr, *h_coef = nullptr;
    uint8_t *tmp_n[4] = {nullptr, nullptr, nullptr, nullptr}, *tmp_m = nullptr;
    uint8_t *omega_pows = nullptr;                                        // omega^i, i < n
    uint8_t *chunks = nullptr, *carries = nullptr, *totals = nullptr, *small = nullptr;
    uint32_t* u32buf[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // rep, unused, rep_rank, unused_rank, tile sums / counters
    uint32_t* count_buf = nullptr;                                        // counting-sort histogram + offsets + tile sums
    uint8_t* eval_chunks = nullptr;                                       // MAX_EVALS x (n / 64) chunk values
    dev::EvalReq* d_reqs = nullptr;
    DevBuf xtab;                                                          // g * omega_ext^i two-level
    uint8_t *vanish_inv = nullptr;
    uint32_t* code = nullptr;
    uint8_t* consts = nullptr;
    EvalTables* d_tab_n = nullptr;
    QuotientArgs* d_qargs = nullptr;
    EvalTables tab_n;
    QuotientArgs qargs;
    std::vector<float> phase_ms;
    // transform lane: the lagrange -> coefficient -> extended-coset transforms of a committed column are queued here as
    // soon as the column is final, so they run under the commitments (and the lookup sorts) of the main lanes instead of
    // in a phase of their own.  Default priority, i.e. below the ctx streams (misc.cu): it takes the SMs the MSM leaves idle.
    cudaStream_t ntt_stream = nullptr;
    cudaEvent_t ev_cols_ready = nullptr, ev_ntt_done = nullptr;
    uint32_t dist_slice = 0, dist_halo = 0;   // several GPUs: rows of the extended domain per rank and the rows around a slice the quotient's rotations reach
                                              // (0: every rank keeps whole extended columns)
    size_t dist_cols_done = 0;   // columns handed to transform_columns so far in this proof (round-robin owner when several GPUs share it)
    bool overlap_ntt = true;
    bool ready = false;   // set at the very end of h2a_circuit_set_keys
};

void h2a_prover_state_free(h2a_ctx* ctx, ProverState* p) {
    if (!p) return;
    cudaStreamSynchronize(ctx->stream);
    if (p->ntt_stream) { cudaStreamSynchronize(p->ntt_stream); cudaStreamDestroy(p->ntt_stream); }
    if (p->ev_cols_ready) cudaEventDestroy(p->ev_cols_ready);
    if (p->ev_ntt_done) cudaEventDestroy(p->ev_ntt_done);
    if (p->arena) cudaFree(p->arena);
    if (p->xtab.p) cudaFree(p->xtab.p);
    delete p;
}

namespace {

uint8_t* arena_take(ProverState* p, size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    uint8_t* r = p->arena + p->used;
    p->used += bytes;
    return r;
}

struct Stepper {  // phase timing with events on the ctx stream
    h2a_ctx* ctx;
    std::vector<cudaEvent_t> ev;
    std::vector<const char*> names;
    void mark(const char* name) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, ctx->stream);
        ev.push_back(e);
        names.push_back(name);
    }
};

#define LAUNCH1D(kernel, count, threads, ...)                                                  \
    do {                                                                                       \
        kernel<<<(unsigned)(((count) + (threads)-1) / (threads)), (threads), 0, ctx->stream>>>(__VA_ARGS__); \
        H2A_LAUNCH_CHECK(ctx);                                                                 \
    } while (0)

int upload_fr(h2a_ctx* ctx, uint8_t* d, const hh::Fr& v) {
    uint8_t b[32];
    hh::fr_store(b, v);
    H2A_CUDA(ctx, cudaMemcpyAsync(d, b, 32, cudaMemcpyHostToDevice, ctx->stream));
    H2A_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // b is a stack buffer
    return H2A_OK;
}

int to_coef(h2a_ctx* ctx, h2a_circuit* c, const uint8_t* lag, uint8_t* coef) {
    uint8_t w[32];
    hh::fr_store(w, c->shape.omega);
    return h2a_ntt_run(ctx, lag, c->shape.n, c->prover->tmp_n[3], coef, c->shape.k, w, 1, nullptr);
}
// One proof over several GPUs through the library's own communicator (h2a_circuit_set_distribution without a callback)
bool native_dist(const h2a_ctx* ctx, const h2a_circuit* c) { return c->dist_world > 1 && !c->dist_exchange && h2a_comm_active(ctx); }

int to_ext(h2a_ctx* ctx, h2a_circuit* c, const uint8_t* coef, uint8_t* ext) {
    uint8_t w[32];
    hh::fr_store(w, hh::fr_root_of_unity((int)c->shape.ext_k));
    return h2a_ntt_run(ctx, coef, c->shape.n, c->prover->tmp_m, ext, c->shape.ext_k, w, 0, c->prover->coset);
}
// to_coef + to_ext of a group of columns on the transform lane, ordered after everything queued on the ctx stream so far
// (the columns' last writes).  tmp_n[3] / tmp_m are the lane's work buffers: nothing else touches them during a proof.
// Without the lane (H2A_PROVE_NTT_OVERLAP=0) the same transforms are queued on the ctx stream itself.
int transform_columns(h2a_ctx* ctx, h2a_circuit* c, const std::vector<Poly3*>& cols, const std::vector<int>* owners = nullptr) {
    ProverState* p = c->prover;
    if (cols.empty()) return H2A_OK;
    // several GPUs: column q of this proof's running count is transformed by rank q % world only, and its coefficient and
    // extended forms are broadcast from there (bulk communicator, same stream as the transforms, one NCCL group per call)
    const bool dist = native_dist(ctx, c);
    const int W = c->dist_world, me = c->dist_rank;
    std::vector<int> owner(cols.size(), me);
    if (dist) for (size_t j = 0; j < cols.size(); j++) owner[j] = owners ? (*owners)[j] : (int)((p->dist_cols_done + j) % (size_t)W);
    if (!owners) p->dist_cols_done += cols.size();
    const uint32_t n = c->shape.n, m = 1u << c->shape.ext_k;
    auto run = [&](cudaStream_t lane) -> int {
        for (size_t j = 0; j < cols.size(); j++)
            if (owner[j] == me) { H2A_TRY(to_coef(ctx, c, cols[j]->lag, cols[j]->coef)); H2A_TRY(to_ext(ctx, c, cols[j]->coef, cols[j]->ext)); }
        if (dist) {
            // the coefficient form goes to everyone (evaluations and openings read whole polynomials); of the extended form a
            // rank needs only the rows its slice of the quotient touches — its m / W rows and `halo` rows on either side, at
            // their own place in the full-size array — so the owner SENDS each rank that window (W times less traffic than a
            // broadcast of 32 m bytes per column)
            H2A_TRY(h2a_comm_group_start(ctx));
            for (size_t j = 0; j < cols.size(); j++) H2A_TRY(h2a_comm_broadcast_on(ctx, 1, cols[j]->coef, 32ull * n, owner[j], lane));
            const uint32_t slice = p->dist_slice, halo = p->dist_halo;
            for (size_t j = 0; j < cols.size(); j++) {
                if (!slice || halo == UINT32_MAX) { H2A_TRY(h2a_comm_broadcast_on(ctx, 1, cols[j]->ext, 32ull * m, owner[j], lane)); continue; }
                for (int r = 0; r < W; r++) {
                    if (r == owner[j] || (me != owner[j] && me != r)) continue;
                    // window of rank r: [r slice - halo, (r + 1) slice + halo) mod m, as one or two contiguous pieces (dist_layout.hpp)
                    DistPiece piece[2];
                    const int pieces = dist_window(m, W, r, halo, piece);
                    uint32_t lo[2] = {piece[0].first, piece[1].first}, len[2] = {piece[0].count, piece[1].count};
                    for (int q = 0; q < pieces; q++) {
                        if (me == owner[j]) H2A_TRY(h2a_comm_send_on(ctx, 1, cols[j]->ext + 32ull * lo[q], 32ull * len[q], r, lane));
                        else H2A_TRY(h2a_comm_recv_on(ctx, 1, cols[j]->ext + 32ull * lo[q], 32ull * len[q], owner[j], lane));
                    }
                }
            }
            H2A_TRY(h2a_comm_group_end(ctx));
        }
        return H2A_OK;
    };
    if (!p->overlap_ntt) return run(ctx->stream);
    H2A_CUDA(ctx, cudaEventRecord(p->ev_cols_ready, ctx->stream));
    H2A_CUDA(ctx, cudaStreamWaitEvent(p->ntt_stream, p->ev_cols_ready, 0));
    struct Swap {   // h2a_ntt_run launches on ctx->stream
        h2a_ctx* c; cudaStream_t saved;
        Swap(h2a_ctx* c_, cudaStream_t s_) : c(c_), saved(c_->stream) { c->stream = s_; }
        ~Swap() { c->stream = saved; }
    } swap(ctx, p->ntt_stream);
    return run(p->ntt_stream);
}
// the ctx stream waits for everything queued on the transform lane
int join_transforms(h2a_ctx* ctx, h2a_circuit* c) {
    ProverState* p = c->prover;
    if (!p->overlap_ntt) return H2A_OK;
    H2A_CUDA(ctx, cudaEventRecord(p->ev_ntt_done, p->ntt_stream));
    H2A_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, p->ev_ntt_done, 0));
    return H2A_OK;
}

int commit(h2a_ctx* ctx, const h2a_bases* bases, const uint8_t* d_scalars, uint32_t n, hh::PointA& out) {
    uint8_t b[64];
    H2A_TRY(h2a_msm_run(ctx, bases, 0, d_scalars, n, b));
    out = hh::affine_load(b);
    return H2A_OK;
}

// commitments of several device-resident columns over the same bases, pipelined over the two MSM lanes
// (host_src, optional: column j is still in host memory there and is copied into cols[j] on the way, overlapped with the
// previous column's MSM)
int commit_batch(h2a_ctx* ctx, const h2a_circuit* c, const h2a_bases* bases, const std::vector<const uint8_t*>& cols, uint32_t n,
                 std::vector<hh::PointA>& out, const std::vector<const uint8_t*>* host_src = nullptr, const std::vector<int>* owners = nullptr) {
    std::vector<uint8_t> pts(64 * cols.size(), 0);
    auto owner_of = [&](size_t j) { return owners ? (*owners)[j] : (int)(j % (size_t)c->dist_world); };   // column j is committed by this rank
    if (c->dist_world > 1 && (c->dist_exchange || h2a_comm_active(ctx))) {   // this rank's share of the columns, then the exchange
        if (host_src) {   // every rank needs every column on its device later
            const bool own_only = native_dist(ctx, c);
            // native distribution: a rank copies only the columns it commits and the others arrive over NVLink (broadcast from
            // their owners) — eight processes copying every column from host memory at once share one host's memory and PCIe
            // paths (k=20, 8 GPUs: 15 ms for this phase against 11 ms on one GPU)
            for (size_t j = 0; j < cols.size(); j++)
                if (!own_only || owner_of(j) == c->dist_rank)
                    H2A_CUDA(ctx, cudaMemcpyAsync((void*)cols[j], (*host_src)[j], 32ull * n, cudaMemcpyHostToDevice, ctx->stream));
            if (own_only) {
                H2A_TRY(h2a_comm_group_start(ctx));
                for (size_t j = 0; j < cols.size(); j++) H2A_TRY(h2a_comm_broadcast_on(ctx, 1, (void*)cols[j], 32ull * n, owner_of(j), ctx->stream));
                H2A_TRY(h2a_comm_group_end(ctx));
            }
        }
        std::vector<const uint8_t*> mine;
        std::vector<size_t> idx;
        for (size_t j = 0; j < cols.size(); j++)
            if (owner_of(j) == c->dist_rank) { mine.push_back(cols[j]); idx.push_back(j); }
        std::vector<size_t> ns(mine.size(), n);
        std::vector<uint8_t> part(64 * mine.size() + 64);
        if (!mine.empty()) H2A_TRY(h2a_msm_batch_dev(ctx, bases, mine.data(), ns.data(), (int)mine.size(), part.data()));
        for (size_t q = 0; q < idx.size(); q++) memcpy(pts.data() + 64 * idx[q], part.data() + 64 * q, 64);
        if (c->dist_exchange) {
            if (c->dist_exchange(c->dist_user, pts.data(), cols.size()) != 0)
                H2A_FAIL(ctx, H2A_ERR_INVALID, "create_proof: the commitment exchange callback failed");
        } else {   // the library's own allgather of raw bytes: slot j comes from its owner, rank j % world
            std::vector<uint8_t> all(pts.size() * (size_t)c->dist_world);
            H2A_TRY(h2a_comm_allgather(ctx, pts.data(), all.data(), pts.size()));
            for (size_t j = 0; j < cols.size(); j++)
                memcpy(pts.data() + 64 * j, all.data() + pts.size() * (size_t)owner_of(j) + 64 * j, 64);
        }
    } else {
        std::vector<size_t> ns(cols.size(), n);
        H2A_TRY(h2a_msm_batch_dev(ctx, bases, cols.data(), ns.data(), (int)cols.size(), pts.data(), host_src ? host_src->data() : nullptr));
    }
    out.resize(cols.size());
    for (size_t i = 0; i < cols.size(); i++) out[i] = hh::affine_load(pts.data() + 64 * i);
    return H2A_OK;
}

// inclusive prefix product of a[0..n) in place (tiles, then recursion over the tile totals)
int scan_mul(h2a_ctx* ctx, uint8_t* a, uint32_t n, uint8_t* totals) {
    const uint32_t tiles = (n + dev::SCAN_TILE - 1) / dev::SCAN_TILE;
    dev::scan_mul_tiles_kernel<<<tiles, dev::SCAN_T, 0, ctx->stream>>>(a, n, totals);
    H2A_LAUNCH_CHECK(ctx);
    if (tiles > 1) {
        H2A_TRY(scan_mul(ctx, totals, tiles, totals + 32ull * ((tiles + 7) & ~7u)));
        LAUNCH1D(dev::scan_mul_apply_kernel, n, 256, a, n, totals);
    }
    return H2A_OK;
}

// ascending bitonic sort of n = 2^log_n 256-bit keys in place
int bitonic_sort(h2a_ctx* ctx, uint8_t* keys, uint32_t n) {
    if (!ctx->sort_attr_set) {
        H2A_CUDA(ctx, cudaFuncSetAttribute(dev::bitonic_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        ctx->sort_attr_set = true;
    }
    for (uint32_t kk = 2; kk <= n; kk <<= 1) {
        uint32_t j = kk >> 1;
        for (; j >= 1024 && n >= 2048; j >>= 1) LAUNCH1D(dev::bitonic_step_kernel, n, 256, keys, n, j, kk);
        if (n >= 2048) {
            if (j) {
                dev::bitonic_tail_kernel<<<n / 2048, 1024, 65536, ctx->stream>>>(keys, n, j, kk);
                H2A_LAUNCH_CHECK(ctx);
            }
        } else {
            for (; j > 0; j >>= 1) LAUNCH1D(dev::bitonic_step_kernel, n, 256, keys, n, j, kk);
        }
    }
    return H2A_OK;
}

int u32_exclusive_scan(h2a_ctx* ctx, const uint32_t* in, uint32_t n, uint32_t* out, uint32_t* sums, uint32_t* total);

// Sorts the u keys at the front of `keys` (n slots, sentinel tail) ascending.  Keys below 2^COUNT_BITS — range tables,
// selector-gated small inputs — take a counting sort (histogram, prefix, expand); anything else the bitonic network.
// count_buf: 2 * (2^COUNT_BITS + 1) + 4096 u32 of scratch; stats: 2 u32.
int sort_keys(h2a_ctx* ctx, uint8_t* keys, uint32_t u, uint32_t n, uint32_t* count_buf, uint32_t* stats) {
    H2A_CUDA(ctx, cudaMemsetAsync(stats, 0, 8, ctx->stream));
    LAUNCH1D(dev::key_stats_kernel, ((u + 31) / 32) * 32, 256, keys, u, stats);
    uint32_t host_stats[2];
    H2A_CUDA(ctx, cudaMemcpyAsync(host_stats, stats, 8, cudaMemcpyDeviceToHost, ctx->stream));
    H2A_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (host_stats[0] != 0 || host_stats[1] >= (1u << dev::COUNT_BITS)) return bitonic_sort(ctx, keys, n);
    const uint32_t bins = host_stats[1] + 1;
    uint32_t* hist = count_buf;
    uint32_t* offsets = count_buf + (1u << dev::COUNT_BITS) + 1;
    uint32_t* sums = offsets + (1u << dev::COUNT_BITS) + 1;
    H2A_CUDA(ctx, cudaMemsetAsync(hist, 0, 4ull * bins, ctx->stream));
    LAUNCH1D(dev::count_keys_kernel, u, 256, keys, u, hist);
    H2A_TRY(u32_exclusive_scan(ctx, hist, bins, offsets, sums, offsets + bins));
    LAUNCH1D(dev::expand_counts_kernel, n, 256, offsets, bins, u, n, keys);
    return H2A_OK;
}

int u32_exclusive_scan(h2a_ctx* ctx, const uint32_t* in, uint32_t n, uint32_t* out, uint32_t* sums, uint32_t* total) {
    const uint32_t tiles = (n + 2047) / 2048;
    dev::u32_tile_sums_kernel<<<tiles, 256, 0, ctx->stream>>>(in, n, sums);
    H2A_LAUNCH_CHECK(ctx);
    dev::u32_scan_sums_kernel<<<1, 1, 0, ctx->stream>>>(sums, tiles, total);
    H2A_LAUNCH_CHECK(ctx);
    dev::u32_tile_apply_kernel<<<tiles, 256, 0, ctx->stream>>>(in, n, sums, out);
    H2A_LAUNCH_CHECK(ctx);
    return H2A_OK;
}

void compress_point(const hh::PointA& p, uint8_t out[32]) {
    if (hh::is_identity(p)) { memset(out, 0, 32); return; }
    uint64_t x[4], y[4];
    hh::fq_to_raw(p.x, x);
    hh::fq_to_raw(p.y, y);
    memcpy(out, x, 32);
    out[31] |= (uint8_t)((y[0] & 1) << 7);
}

hh::Fr rotate_point(const Shape& s, const hh::Fr& x, int32_t rot) {
    return x * (rot >= 0 ? hh::pow_u64(s.omega, (uint64_t)rot) : hh::pow_u64(s.omega_inv, (uint64_t)(-(int64_t)rot)));
}

}  // namespace

extern "C" {

// Loads the proving key: commits the fixed columns and the permutation polynomials (so the vk is set too) and
// keeps their coefficient and extended-coset forms resident.
static int set_keys_impl(h2a_ctx* ctx, h2a_circuit* c, const h2a_bases* g, const h2a_bases* g_lagrange, const uint8_t* fixed_values,
                         const uint8_t* sigmas, const uint8_t vk_hash[32], const uint8_t coset_shift[32]);
int h2a_circuit_set_keys(h2a_ctx* ctx, h2a_circuit* c, const h2a_bases* g, const h2a_bases* g_lagrange, const uint8_t* fixed_values,
                         const uint8_t* sigmas, const uint8_t vk_hash[32], const uint8_t coset_shift[32]) {
    H2A_DEVICE(ctx);
    const int rc = set_keys_impl(ctx, c, g, g_lagrange, fixed_values, sigmas, vk_hash, coset_shift);
    if (rc != H2A_OK && ctx && c && c->prover) {   // never leave a half-built proving key behind: create_proof would run on it
        h2a_prover_state_free(ctx, c->prover);
        c->prover = nullptr;
    }
    return rc;
}
static int set_keys_impl(h2a_ctx* ctx, h2a_circuit* c, const h2a_bases* g, const h2a_bases* g_lagrange, const uint8_t* fixed_values,
                         const uint8_t* sigmas, const uint8_t vk_hash[32], const uint8_t coset_shift[32]) {
    if (!ctx || !c || !g || !g_lagrange || !vk_hash || !coset_shift) return H2A_ERR_INVALID;
    const Shape& s = c->shape;
    const uint32_t n = s.n, m = 1u << s.ext_k;
    if ((s.n_fixed && !fixed_values) || (!s.perm.empty() && !sigmas)) H2A_FAIL(ctx, H2A_ERR_INVALID, "set_keys: fixed / permutation columns missing");
    if (g->n < n || g_lagrange->n < n) H2A_FAIL(ctx, H2A_ERR_INVALID, "set_keys: params hold fewer than %u bases", n);
    if (s.n_advice > MAX_COLS || s.n_fixed > MAX_COLS || s.n_instance > MAX_COLS || s.aq.size() > MAX_Q || s.fq.size() > MAX_Q ||
        s.iq.size() > MAX_Q || s.lookups.size() > MAX_LK || s.perm.size() > MAX_PERM || s.n_chunks > MAX_CHUNKS || s.chunk_len > 8)
        H2A_FAIL(ctx, H2A_ERR_INVALID, "set_keys: circuit exceeds the prover's column/query limits");
    if (c->prover) h2a_prover_state_free(ctx, c->prover);
    ProverState* p = c->prover = new ProverState();
    p->g = g;
    p->g_lagrange = g_lagrange;
    memcpy(p->coset, coset_shift, 32);
    if (const char* env = getenv("H2A_PROVE_NTT_OVERLAP")) p->overlap_ntt = atoi(env) != 0;
    if (p->overlap_ntt) {
        const char* prio = getenv("H2A_PROVE_NTT_PRIO");   // experiment knob: 1 = same priority as the commitment lanes
        H2A_CUDA(ctx, cudaStreamCreateWithPriority(&p->ntt_stream, cudaStreamNonBlocking, prio && atoi(prio) ? ctx->stream_priority : 0));
        H2A_CUDA(ctx, cudaEventCreateWithFlags(&p->ev_cols_ready, cudaEventDisableTiming));
        H2A_CUDA(ctx, cudaEventCreateWithFlags(&p->ev_ntt_done, cudaEventDisableTiming));
    }

    // programs: gates first, then per lookup its inputs then its tables
    std::vector<uint32_t> code;
    std::vector<uint32_t> off, len;
    auto push = [&](const h2a_plonk::Prog& g_) { off.push_back((uint32_t)code.size() / 2); len.push_back((uint32_t)g_.code.size() / 2); code.insert(code.end(), g_.code.begin(), g_.code.end()); };
    for (auto& g_ : s.gates) push(g_);
    std::vector<uint32_t> in_first, tab_first;
    for (auto& l : s.lookups) {
        in_first.push_back((uint32_t)off.size());
        for (auto& g_ : l.inputs) push(g_);
        tab_first.push_back((uint32_t)off.size());
        for (auto& g_ : l.tables) push(g_);
    }
    if (off.size() > MAX_PROGS) H2A_FAIL(ctx, H2A_ERR_INVALID, "set_keys: more than %d expression programs", MAX_PROGS);

    // arena
    const size_t nl = s.lookups.size();
    const size_t n_arrays = 2 * (s.n_fixed + s.perm.size() + s.n_advice + s.n_instance + s.n_chunks) + nl * (2 + 6) + 1 + 4 + 1 + 8;
    const size_t m_arrays = (s.n_fixed + s.perm.size() + s.n_advice + s.n_instance + s.n_chunks) + nl * 3 + 3 + 2 + 1;
    p->arena_bytes = 32ull * MAX_EVALS * (n / dev::HORNER_L + 1) + sizeof(dev::EvalReq) * MAX_EVALS + 8192 + 5 * (4ull * n + 8192) + 4ull * (2 * ((1u << dev::COUNT_BITS) + 1) + 8192) + n_arrays * (32ull * n + 256) + m_arrays * (32ull * m + 256) + (1 << 20) + code.size() * 4 + 64 * s.consts.size() +
                     sizeof(EvalTables) + sizeof(QuotientArgs);
    H2A_CUDA(ctx, cudaMalloc(&p->arena, p->arena_bytes));
    auto poly3 = [&]() { Poly3 q; q.lag = arena_take(p, 32ull * n); q.coef = arena_take(p, 32ull * n); q.ext = arena_take(p, 32ull * m); return q; };
    for (uint32_t i = 0; i < s.n_fixed; i++) p->fixed.push_back(poly3());
    for (size_t i = 0; i < s.perm.size(); i++) p->sigma.push_back(poly3());
    for (uint32_t i = 0; i < s.n_advice; i++) p->advice.push_back(poly3());
    for (uint32_t i = 0; i < s.n_instance; i++) p->instance.push_back(poly3());
    for (uint32_t i = 0; i < s.n_chunks; i++) p->pz.push_back(poly3());
    for (size_t i = 0; i < nl; i++) {
        ProverState::Lk l;
        l.A = arena_take(p, 32ull * n);
        l.S = arena_take(p, 32ull * n);
        l.pa = poly3(); l.ps = poly3(); l.z = poly3();
        p->lk.push_back(l);
    }
    p->l0 = arena_take(p, 32ull * m); p->llast = arena_take(p, 32ull * m); p->lblind = arena_take(p, 32ull * m);
    p->random_coef = arena_take(p, 32ull * n);
    p->h_ext = arena_take(p, 32ull * m);
    p->h_coef = arena_take(p, 32ull * m);
    for (int i = 0; i < 4; i++) p->tmp_n[i] = arena_take(p, 32ull * n);
    p->tmp_m = arena_take(p, 32ull * m);
    p->omega_pows = arena_take(p, 32ull * n);
    p->chunks = arena_take(p, 32ull * (m / dev::HORNER_L + 64));
    p->carries = arena_take(p, 32ull * (m / dev::HORNER_L + 64));
    p->totals = arena_take(p, 32ull * (n / dev::SCAN_TILE + 4096));
    p->small = arena_take(p, 32 * 1024);
    for (int i = 0; i < 5; i++) p->u32buf[i] = (uint32_t*)arena_take(p, 4ull * n + 4096);
    p->count_buf = (uint32_t*)arena_take(p, 4ull * (2 * ((1u << dev::COUNT_BITS) + 1) + 4096));
    p->eval_chunks = arena_take(p, 32ull * MAX_EVALS * (n / dev::HORNER_L + 1));
    p->d_reqs = (dev::EvalReq*)arena_take(p, sizeof(dev::EvalReq) * MAX_EVALS);
    p->vanish_inv = arena_take(p, 32ull * (m / n));
    p->code = (uint32_t*)arena_take(p, code.size() * 4 + 16);
    p->consts = arena_take(p, 32 * s.consts.size() + 32);
    p->d_tab_n = (EvalTables*)arena_take(p, sizeof(EvalTables));
    p->d_qargs = (QuotientArgs*)arena_take(p, sizeof(QuotientArgs));
    if (p->used > p->arena_bytes) H2A_FAIL(ctx, H2A_ERR_OOM, "set_keys: arena estimate too small (%zu > %zu)", p->used, p->arena_bytes);

    if (!code.empty()) H2A_CUDA(ctx, cudaMemcpyAsync(p->code, code.data(), code.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<uint8_t> cb(32 * s.consts.size() + 32, 0);
    for (size_t i = 0; i < s.consts.size(); i++) hh::fr_store(cb.data() + 32 * i, s.consts[i]);
    H2A_CUDA(ctx, cudaMemcpyAsync(p->consts, cb.data(), cb.size(), cudaMemcpyHostToDevice, ctx->stream));
    H2A_CUDA(ctx, cudaStreamSynchronize(ctx->stream));

    // omega^i and X_i = g * omega_ext^i
    uint8_t wb[32], oneb[32];
    hh::fr_store(oneb, hh::fr_one());
    hh::fr_store(wb, s.omega);
    H2A_TRY(h2a_pow_vector(ctx, wb, oneb, n, p->omega_pows));
    hh::Fr w_ext = hh::fr_root_of_unity((int)s.ext_k), gshift = hh::fr_load(coset_shift);
    hh::fr_store(wb, w_ext);
    H2A_TRY(h2a_pow_tables(ctx, wb, coset_shift, m, p->xtab));
    {   // 1 / (X_i^n - 1) has period m / n
        std::vector<uint8_t> vb(32ull * (m / n));
        hh::Fr gn = gshift, wn = w_ext;
        for (uint32_t i = 0; i < s.k; i++) { gn = hh::sqr(gn); wn = hh::sqr(wn); }
        hh::Fr cur = gn;
        for (uint32_t j = 0; j < m / n; j++) { hh::fr_store(vb.data() + 32 * j, hh::inv(cur - hh::fr_one())); cur = cur * wn; }
        H2A_CUDA(ctx, cudaMemcpyAsync(p->vanish_inv, vb.data(), vb.size(), cudaMemcpyHostToDevice, ctx->stream));
        H2A_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    // fixed + sigma: values, commitments, coefficient and extended forms
    c->fixed_comms.assign(64 * (size_t)s.n_fixed, 0);
    c->sigma_comms.assign(64 * s.perm.size(), 0);
    for (uint32_t i = 0; i < s.n_fixed + s.perm.size(); i++) {
        const bool is_fixed = i < s.n_fixed;
        Poly3& q = is_fixed ? p->fixed[i] : p->sigma[i - s.n_fixed];
        const uint8_t* src = is_fixed ? fixed_values + 32ull * n * i : sigmas + 32ull * n * (i - s.n_fixed);
        H2A_CUDA(ctx, cudaMemcpyAsync(q.lag, src, 32ull * n, cudaMemcpyHostToDevice, ctx->stream));
        hh::PointA cm;
        H2A_TRY(commit(ctx, g_lagrange, q.lag, n, cm));
        hh::affine_store((is_fixed ? c->fixed_comms.data() + 64 * i : c->sigma_comms.data() + 64 * (i - s.n_fixed)), cm);
        H2A_TRY(to_coef(ctx, c, q.lag, q.coef));
        H2A_TRY(to_ext(ctx, c, q.coef, q.ext));
    }
    {   // l_0, l_last, l_blind on the extended coset
        std::vector<uint8_t> lag(32ull * n, 0);
        uint8_t* dst[3] = {p->l0, p->llast, p->lblind};
        for (int which = 0; which < 3; which++) {
            std::fill(lag.begin(), lag.end(), 0);
            if (which == 0) hh::fr_store(lag.data(), hh::fr_one());
            if (which == 1) hh::fr_store(lag.data() + 32ull * s.usable, hh::fr_one());
            if (which == 2) for (uint32_t r = s.usable + 1; r < n; r++) hh::fr_store(lag.data() + 32ull * r, hh::fr_one());
            H2A_CUDA(ctx, cudaMemcpyAsync(p->tmp_n[0], lag.data(), lag.size(), cudaMemcpyHostToDevice, ctx->stream));
            H2A_TRY(to_coef(ctx, c, p->tmp_n[0], p->tmp_n[1]));
            H2A_TRY(to_ext(ctx, c, p->tmp_n[1], dst[which]));
            H2A_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
    }
    memcpy(c->vk_hash, vk_hash, 32);
    c->has_vk = true;

    // evaluation tables
    auto fill_tables = [&](EvalTables& t, bool ext) {
        memset(&t, 0, sizeof t);
        for (uint32_t i = 0; i < s.n_advice; i++) t.col[0][i] = ext ? p->advice[i].ext : p->advice[i].lag;
        for (uint32_t i = 0; i < s.n_fixed; i++) t.col[1][i] = ext ? p->fixed[i].ext : p->fixed[i].lag;
        for (uint32_t i = 0; i < s.n_instance; i++) t.col[2][i] = ext ? p->instance[i].ext : p->instance[i].lag;
        const std::vector<h2a_plonk::Query>* qs[3] = {&s.aq, &s.fq, &s.iq};
        for (int ty = 0; ty < 3; ty++)
            for (size_t i = 0; i < qs[ty]->size(); i++) { t.q_col[ty][i] = (*qs[ty])[i].col; t.q_rot[ty][i] = (*qs[ty])[i].rot; }
        t.consts = p->consts;
        t.code = p->code;
        for (size_t i = 0; i < off.size(); i++) { t.prog_off[i] = off[i]; t.prog_len[i] = len[i]; }
        t.mask = (ext ? m : n) - 1;
        t.step = ext ? m / n : 1;
    };
    fill_tables(p->tab_n, false);
    QuotientArgs& qa = p->qargs;
    memset(&qa, 0, sizeof qa);
    fill_tables(qa.t, true);
    qa.n_gates = (uint32_t)s.gates.size();
    qa.n_lookups = (uint32_t)nl;
    for (size_t i = 0; i < nl; i++) {
        qa.lk_in_first[i] = in_first[i]; qa.lk_in_cnt[i] = (uint32_t)s.lookups[i].inputs.size();
        qa.lk_tab_first[i] = tab_first[i]; qa.lk_tab_cnt[i] = (uint32_t)s.lookups[i].tables.size();
        qa.lk_pa[i] = p->lk[i].pa.ext; qa.lk_ps[i] = p->lk[i].ps.ext; qa.lk_z[i] = p->lk[i].z.ext;
    }
    qa.n_perm = (uint32_t)s.perm.size(); qa.chunk_len = s.chunk_len; qa.n_chunks = s.n_chunks;
    for (size_t i = 0; i < s.perm.size(); i++) { qa.perm_type[i] = s.perm[i].type; qa.perm_qidx[i] = s.perm[i].qidx; qa.sigma[i] = p->sigma[i].ext; }
    for (uint32_t i = 0; i < s.n_chunks; i++) qa.pz[i] = p->pz[i].ext;
    qa.l0 = p->l0; qa.llast = p->llast; qa.lblind = p->lblind;
    qa.x_lo = (const uint8_t*)p->xtab.p; qa.x_hi = qa.x_lo + (32ull << LOG_PW);
    qa.vanish_inv = p->vanish_inv; qa.vanish_mask = m / n - 1;
    qa.last_rot = s.last_rot;
    qa.out = p->h_ext;
    H2A_CUDA(ctx, cudaMemcpyAsync(p->d_tab_n, &p->tab_n, sizeof(EvalTables), cudaMemcpyHostToDevice, ctx->stream));
    H2A_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    p->ready = true;
    return H2A_OK;
}

int h2a_circuit_set_distribution(h2a_ctx* ctx, h2a_circuit* c, int rank, int world, h2a_exchange_fn exchange, void* user) {
    if (!ctx || !c || world < 1 || rank < 0 || rank >= world) return H2A_ERR_INVALID;
    if (world > 1 && !exchange && (!h2a_comm_active(ctx) || ctx->comm_world != world || ctx->comm_rank != rank))
        H2A_FAIL(ctx, H2A_ERR_INVALID, "set_distribution: no exchange callback and no communicator of %d ranks with this rank = %d (h2a_comm_init)", world, rank);
    c->dist_rank = rank;
    c->dist_world = world;
    c->dist_exchange = exchange;
    c->dist_user = user;
    return H2A_OK;
}

int h2a_circuit_get_vk(h2a_ctx* ctx, const h2a_circuit* c, uint8_t* fixed_comms, uint8_t* sigma_comms) {
    H2A_DEVICE(ctx);
    if (!ctx || !c) return H2A_ERR_INVALID;
    if (!c->has_vk) H2A_FAIL(ctx, H2A_ERR_INVALID, "get_vk: no key set");
    if (fixed_comms && !c->fixed_comms.empty()) memcpy(fixed_comms, c->fixed_comms.data(), c->fixed_comms.size());
    if (sigma_comms && !c->sigma_comms.empty()) memcpy(sigma_comms, c->sigma_comms.data(), c->sigma_comms.size());
    return H2A_OK;
}

size_t h2a_blinds_len(const h2a_circuit* c) {
    if (!c) return 0;
    const Shape& s = c->shape;
    return s.lookups.size() * (2 * (size_t)(s.bf + 1) + s.bf) + (size_t)s.n_chunks * s.bf + s.n;
}

size_t h2a_proof_len(const h2a_circuit* c) {
    if (!c) return 0;
    const Shape& s = c->shape;
    std::map<int32_t, int> rots;
    for (auto& q : s.iq) rots[q.rot] = 1;
    for (auto& q : s.aq) rots[q.rot] = 1;
    for (auto& q : s.fq) rots[q.rot] = 1;
    rots[0] = 1;
    if (s.n_chunks) rots[1] = 1;
    if (s.n_chunks > 1) rots[s.last_rot] = 1;
    if (!s.lookups.empty()) { rots[1] = 1; rots[-1] = 1; }
    size_t points = s.n_advice + 3 * s.lookups.size() + s.n_chunks + 1 + s.qdeg + rots.size();
    size_t scalars = s.iq.size() + s.aq.size() + s.fq.size() + 1 + s.perm.size() + (s.n_chunks ? 3 * (size_t)s.n_chunks - 1 : 0) +
                     5 * s.lookups.size();
    return 32 * (points + scalars);
}

int h2a_create_proof(h2a_ctx* ctx, h2a_circuit* c, const uint8_t* instance_cols, const uint8_t* advice_cols, const uint8_t* blinds,
                     uint8_t* proof_out, size_t proof_cap, size_t* proof_len, uint8_t* inst_comms_out) {
    H2A_DEVICE(ctx);
    if (!ctx || !c || !proof_out || !proof_len || !blinds) return H2A_ERR_INVALID;
    ProverState* p = c->prover;
    if (!p || !p->ready) H2A_FAIL(ctx, H2A_ERR_INVALID, "create_proof: no proving key (h2a_circuit_set_keys)");
    const Shape& s = c->shape;
    const uint32_t n = s.n, m = 1u << s.ext_k, u = s.usable, bf = s.bf;
    if (proof_cap < h2a_proof_len(c)) H2A_FAIL(ctx, H2A_ERR_INVALID, "create_proof: output buffer too small");
    if ((s.n_instance && !instance_cols) || (s.n_advice && !advice_cols)) return H2A_ERR_INVALID;
    cudaStream_t st = ctx->stream;
    // the per-call MSM / NTT phase events (h2a_set_profiling) end in a stream synchronisation each: inside a proof that
    // would serialise the transform lane against the host, and the proof has its own phase events (Stepper)
    struct ProfOff {
        h2a_ctx* c; bool saved;
        explicit ProfOff(h2a_ctx* c_) : c(c_), saved(c_->profiling) { c->profiling = false; }
        ~ProfOff() { c->profiling = saved; }
    } prof_off(ctx);
    Stepper steps{ctx};
    struct Guard {   // on EVERY exit path: nothing of this proof stays in flight on the transform lane (an immediate retry would race
                     // it on the column buffers and tmp_n[3] / tmp_m), and the phase events are released
        h2a_ctx* ctx; ProverState* p; Stepper* steps;
        ~Guard() {
            if (p->ntt_stream) cudaStreamSynchronize(p->ntt_stream);
            if (ctx->alt) cudaStreamSynchronize(ctx->alt->stream);
            for (cudaEvent_t e : steps->ev) cudaEventDestroy(e);
            steps->ev.clear();
        }
    } guard{ctx, p, &steps};
    steps.mark("start");
    p->dist_cols_done = 0;
    p->dist_slice = p->dist_halo = 0;
    if (native_dist(ctx, c) && m % (uint32_t)c->dist_world == 0 && m / (uint32_t)c->dist_world >= 1024) {
        // rows a quotient row reaches: every query rotation, +-1 (grand products, lookups) and the permutation's last_rot,
        // each times the step m / n of the extended domain
        int64_t reach = 1;
        for (auto* qs : {&s.aq, &s.fq, &s.iq}) for (auto& q : *qs) reach = std::max<int64_t>(reach, std::llabs((long long)q.rot));
        reach = std::max<int64_t>(reach, std::llabs((long long)s.last_rot));
        const uint64_t halo = (uint64_t)reach * (m / n);
        p->dist_slice = m / (uint32_t)c->dist_world;
        if (2 * halo <= p->dist_slice) p->dist_halo = (uint32_t)halo;
        else p->dist_halo = UINT32_MAX;   // a reach wider than half a slice: quotient rows still split, whole extended columns broadcast
    }

    h2a_glue::Transcript tr;
    size_t pos = 0;
    // a commitment the transcript absorbs must not be the identity: the verifier's read_point rejects it (and common_point skips
    // it, so the two transcripts would part ways); returns false then.  The final W_i are written without being absorbed,
    // as the verifier reads them (src/multiopen.rs:202-218).
    auto write_point = [&](const hh::PointA& pt) -> bool {
        const bool ok = tr.common_point(pt);
        compress_point(pt, proof_out + pos);
        pos += 32;
        return ok;
    };
    auto write_points = [&](const std::vector<hh::PointA>& pts, const char* what) -> int {
        for (size_t i = 0; i < pts.size(); i++)
            if (!write_point(pts[i])) H2A_FAIL(ctx, H2A_ERR_INVALID, "create_proof: %s commitment %zu is the identity (e.g. an all-zero unblinded column)", what, i);
        return H2A_OK;
    };
    auto write_scalar = [&](const hh::Fr& v) { tr.common_scalar(v); uint64_t raw[4]; hh::fr_to_raw(v, raw); memcpy(proof_out + pos, raw, 32); pos += 32; };
    const uint8_t* bl = blinds;

    // small device scalars: slots of 32 bytes in p->small
    auto slot = [&](int i) { return p->small + 32 * i; };
    enum { S_THETA = 0, S_BETA, S_GAMMA, S_Y, S_V, S_INIT, S_Z0, S_EVAL0 = 16 };

    tr.common_scalar(hh::fr_load(c->vk_hash));                                     // src/verifier.rs:341-358
    {   // instance (:360-363) and advice (:365-376) commitments in one pipelined batch
        std::vector<const uint8_t*> cols, src;
        for (uint32_t i = 0; i < s.n_instance; i++) {
            src.push_back(instance_cols + 32ull * n * i);
            cols.push_back(p->instance[i].lag);
        }
        for (uint32_t i = 0; i < s.n_advice; i++) {
            src.push_back(advice_cols + 32ull * n * i);
            cols.push_back(p->advice[i].lag);
        }
        std::vector<hh::PointA> cms;
        H2A_TRY(commit_batch(ctx, c, p->g_lagrange, cols, n, cms, &src));
        for (uint32_t i = 0; i < s.n_instance; i++) {
            if (!tr.common_point(cms[i])) H2A_FAIL(ctx, H2A_ERR_INVALID, "create_proof: instance column %u commits to the identity", i);
            if (inst_comms_out) hh::affine_store(inst_comms_out + 64 * i, cms[i]);
        }
        for (uint32_t i = 0; i < s.n_advice; i++)
            if (!write_point(cms[s.n_instance + i])) H2A_FAIL(ctx, H2A_ERR_INVALID, "create_proof: advice column %u commits to the identity", i);
    }
    {   // the columns are resident now: their transforms run under the lookup permutations below
        std::vector<Poly3*> cols;
        for (auto& q : p->advice) cols.push_back(&q);
        for (auto& q : p->instance) cols.push_back(&q);
        H2A_TRY(transform_columns(ctx, c, cols));
    }
    steps.mark("instance+advice commitments");
    hh::Fr theta = tr.squeeze();                                                   // :378
    H2A_TRY(upload_fr(ctx, slot(S_THETA), theta));

    // several GPUs (native distribution): lookup li belongs to rank li % world from its permutation to its grand product —
    // only the owner sorts it, commits A', S', Z and transforms them; the other ranks receive the commitments and the
    // coefficient / extended forms.  Permutation chunks are computed everywhere (cheap) and committed round-robin.
    const bool dist = native_dist(ctx, c);
    const int dist_me = c->dist_rank, dist_w = dist ? c->dist_world : 1;
    auto lk_owner = [&](size_t li) { return dist ? (int)(li % (size_t)dist_w) : dist_me; };
    auto pz_owner = [&](size_t ci) { return dist ? (int)((ci + 1 + s.lookups.size()) % (size_t)dist_w) : dist_me; };
    uint32_t lookup_err = 0;                                                       // 1 + index of a lookup whose input is not in its table
    for (size_t li = 0; li < s.lookups.size(); li++) {                             // :380-387, src/lookup.rs:49-79
        ProverState::Lk& l = p->lk[li];
        if (lk_owner(li) != dist_me) { bl += 64ull * (n - u) + 32ull * bf; continue; }
        LAUNCH1D(dev::compress_kernel, n, 128, p->d_tab_n, p->qargs.lk_in_first[li], p->qargs.lk_in_cnt[li], slot(S_THETA), n, l.A);
        LAUNCH1D(dev::compress_kernel, n, 128, p->d_tab_n, p->qargs.lk_tab_first[li], p->qargs.lk_tab_cnt[li], slot(S_THETA), n, l.S);
        {   // A' and S' on the device: sort both, match runs to table entries, hand out the left-overs
            uint8_t *ka = p->tmp_n[0], *ks = p->tmp_n[1], *leftover = p->tmp_n[2];
            uint32_t *rep = p->u32buf[0], *unused = p->u32buf[1], *rep_rank = p->u32buf[2], *unused_rank = p->u32buf[3];
            uint32_t *sums = p->u32buf[4], *counters = p->u32buf[4] + (n / 2048 + 8);   // counters: [0] n_rep, [1] n_unused, [2] error
            LAUNCH1D(dev::lookup_keys_kernel, n, 256, l.A, u, n, ka);
            LAUNCH1D(dev::lookup_keys_kernel, n, 256, l.S, u, n, ks);
            H2A_TRY(sort_keys(ctx, ka, u, n, p->count_buf, counters + 4));
            H2A_TRY(sort_keys(ctx, ks, u, n, p->count_buf, counters + 4));
            LAUNCH1D(dev::fill_u32_kernel, u, 256, unused, u, 1u);
            H2A_CUDA(ctx, cudaMemsetAsync(counters, 0, 16, st));
            LAUNCH1D(dev::lookup_match_kernel, u, 128, ka, ks, u, rep, unused, counters + 2);
            H2A_TRY(u32_exclusive_scan(ctx, rep, u, rep_rank, sums, counters + 0));
            H2A_TRY(u32_exclusive_scan(ctx, unused, u, unused_rank, sums, counters + 1));
            LAUNCH1D(dev::lookup_leftover_kernel, u, 256, ks, unused, unused_rank, u, leftover);
            LAUNCH1D(dev::lookup_assign_kernel, u, 128, ka, leftover, rep, rep_rank, counters + 0, u, l.pa.lag, l.ps.lag);
            uint32_t host_counters[4];
            H2A_CUDA(ctx, cudaMemcpyAsync(host_counters, counters, 16, cudaMemcpyDeviceToHost, st));
            H2A_CUDA(ctx, cudaStreamSynchronize(st));
            if ((host_counters[2] || host_counters[0] != host_counters[1]) && !lookup_err) lookup_err = (uint32_t)li + 1;
        }
        H2A_CUDA(ctx, cudaMemcpyAsync(l.pa.lag + 32ull * u, bl, 32ull * (n - u), cudaMemcpyHostToDevice, st));
        bl += 32ull * (n - u);
        H2A_CUDA(ctx, cudaMemcpyAsync(l.ps.lag + 32ull * u, bl, 32ull * (n - u), cudaMemcpyHostToDevice, st));
        bl += 32ull * (n - u);
        bl += 32ull * bf;  // this lookup's Z tail, consumed after beta and gamma
    }
    if (dist && !s.lookups.empty()) {   // every rank must learn of a failed lookup before any of them enters the next collective
        std::vector<uint8_t> all(4 * (size_t)dist_w);
        H2A_TRY(h2a_comm_allgather(ctx, (const uint8_t*)&lookup_err, all.data(), 4));
        for (int r = 0; r < dist_w && !lookup_err; r++) memcpy(&lookup_err, all.data() + 4 * r, 4);
    }
    if (lookup_err) H2A_FAIL(ctx, H2A_ERR_INVALID, "create_proof: lookup %u has an input value absent from its table", lookup_err - 1);
    if (!s.lookups.empty()) {
        std::vector<int> owners;
        for (size_t li = 0; li < s.lookups.size(); li++) { owners.push_back(lk_owner(li)); owners.push_back(lk_owner(li)); }
        {
            std::vector<Poly3*> tcols;
            for (auto& l : p->lk) { tcols.push_back(&l.pa); tcols.push_back(&l.ps); }
            H2A_TRY(transform_columns(ctx, c, tcols, dist ? &owners : nullptr));
        }
        std::vector<const uint8_t*> cols;
        for (auto& l : p->lk) { cols.push_back(l.pa.lag); cols.push_back(l.ps.lag); }
        std::vector<hh::PointA> cms;
        H2A_TRY(commit_batch(ctx, c, p->g_lagrange, cols, n, cms, nullptr, dist ? &owners : nullptr));
        H2A_TRY(write_points(cms, "permuted lookup column"));
    }
    steps.mark("lookup permuted columns");
    hh::Fr beta = tr.squeeze(), gamma = tr.squeeze();                              // :390,393
    H2A_TRY(upload_fr(ctx, slot(S_BETA), beta));
    H2A_TRY(upload_fr(ctx, slot(S_GAMMA), gamma));

    static const uint64_t DELTA_RAW[4] = {0x870e56bbe533e9a2ull, 0x5b5f898e5e963f25ull, 0x64ec26aad4c86e71ull, 0x09226b6e22c6f0caull};
    const hh::Fr delta = hh::fr_from_raw(DELTA_RAW);
    {   // beta * delta^i for the quotient kernel
        hh::Fr bd = beta;
        for (size_t i = 0; i < s.perm.size(); i++) { hh::fr_store(p->qargs.beta_delta[i], bd); bd = bd * delta; }
    }
    auto column_lag = [&](const h2a_plonk::PermCol& pc) {
        return pc.type == h2a_plonk::COL_ADVICE ? p->advice[pc.col].lag : pc.type == h2a_plonk::COL_FIXED ? p->fixed[pc.col].lag : p->instance[pc.col].lag;
    };
    // blinds of the lookup grand products come before the permutation ones in the buffer: remember the cursor
    const uint8_t* bl_lookup_z[MAX_LK];
    // layout: per lookup (A' tail, S' tail, Z tail) — Z tails were skipped above; recompute cursors explicitly
    {
        const uint8_t* b = blinds;
        for (size_t li = 0; li < s.lookups.size(); li++) { b += 64ull * (n - u); bl_lookup_z[li] = b; b += 32ull * bf; }
        bl = b;  // permutation blinds start here
    }
    H2A_TRY(upload_fr(ctx, slot(S_INIT), hh::fr_one()));
    for (uint32_t ci = 0; ci < s.n_chunks; ci++) {                                 // :402-409, src/permutation.rs:48-78
        dev::PermTermArgs a;
        memset(&a, 0, sizeof a);
        const uint32_t lo = ci * s.chunk_len, hi = (uint32_t)std::min<size_t>((ci + 1) * s.chunk_len, s.perm.size());
        a.ncols = hi - lo;
        for (uint32_t k = lo; k < hi; k++) {
            a.vals[k - lo] = column_lag(s.perm[k]);
            a.sigma[k - lo] = p->sigma[k].lag;
            memcpy(a.beta_delta[k - lo], p->qargs.beta_delta[k], 32);
        }
        LAUNCH1D(dev::perm_terms_kernel, n, 128, a, slot(S_BETA), slot(S_GAMMA), p->omega_pows, n, p->tmp_n[0], p->tmp_n[1]);
        LAUNCH1D(dev::batch_inverse_kernel, (n + dev::INV_CHUNK - 1) / dev::INV_CHUNK, 128, p->tmp_n[1], n);
        LAUNCH1D(dev::mul_arrays_kernel, n, 256, p->tmp_n[0], p->tmp_n[1], n, p->tmp_n[2]);
        H2A_TRY(scan_mul(ctx, p->tmp_n[2], n, p->totals));
        LAUNCH1D(dev::shift_scale_kernel, n, 256, p->tmp_n[2], slot(S_INIT), n, p->pz[ci].lag);
        // last_z of this chunk (row `usable`) seeds the next chunk, read before the blinds overwrite the tail
        H2A_CUDA(ctx, cudaMemcpyAsync(slot(S_INIT), p->pz[ci].lag + 32ull * u, 32, cudaMemcpyDeviceToDevice, st));
        H2A_CUDA(ctx, cudaMemcpyAsync(p->pz[ci].lag + 32ull * (n - bf), bl, 32ull * bf, cudaMemcpyHostToDevice, st));
        bl += 32ull * bf;
    }
    steps.mark("permutation grand products");
    for (size_t li = 0; li < s.lookups.size(); li++) {                             // :411-417, src/lookup.rs:81-106
        ProverState::Lk& l = p->lk[li];
        if (lk_owner(li) != dist_me) continue;
        LAUNCH1D(dev::lookup_terms_kernel, n, 128, l.A, l.S, l.pa.lag, l.ps.lag, slot(S_BETA), slot(S_GAMMA), u, n, p->tmp_n[0], p->tmp_n[1]);
        LAUNCH1D(dev::batch_inverse_kernel, (n + dev::INV_CHUNK - 1) / dev::INV_CHUNK, 128, p->tmp_n[1], n);
        LAUNCH1D(dev::mul_arrays_kernel, n, 256, p->tmp_n[0], p->tmp_n[1], n, p->tmp_n[2]);
        H2A_TRY(scan_mul(ctx, p->tmp_n[2], n, p->totals));
        H2A_TRY(upload_fr(ctx, slot(S_Z0), hh::fr_one()));
        LAUNCH1D(dev::shift_scale_kernel, n, 256, p->tmp_n[2], slot(S_Z0), n, l.z.lag);
        H2A_CUDA(ctx, cudaMemcpyAsync(l.z.lag + 32ull * (n - bf), bl_lookup_z[li], 32ull * bf, cudaMemcpyHostToDevice, st));
    }
    {
        std::vector<int> owners;
        for (size_t ci = 0; ci < p->pz.size(); ci++) owners.push_back(pz_owner(ci));
        for (size_t li = 0; li < p->lk.size(); li++) owners.push_back(lk_owner(li));
        {
            std::vector<Poly3*> tcols;
            for (auto& q : p->pz) tcols.push_back(&q);
            for (auto& l : p->lk) tcols.push_back(&l.z);
            H2A_TRY(transform_columns(ctx, c, tcols, dist ? &owners : nullptr));
        }
        // permutation Z (:402-409) then lookup Z (:411-417) commitments, one batch
        std::vector<const uint8_t*> cols;
        for (auto& q : p->pz) cols.push_back(q.lag);
        for (auto& l : p->lk) cols.push_back(l.z.lag);
        std::vector<hh::PointA> cms;
        H2A_TRY(commit_batch(ctx, c, p->g_lagrange, cols, n, cms, nullptr, dist ? &owners : nullptr));
        H2A_TRY(write_points(cms, "grand product"));
    }
    steps.mark("lookup grand products + Z commitments");
    H2A_CUDA(ctx, cudaMemcpyAsync(p->random_coef, bl, 32ull * n, cudaMemcpyHostToDevice, st));   // src/vanishing.rs:54-75
    {
        hh::PointA cm;
        H2A_TRY(commit(ctx, p->g, p->random_coef, n, cm));
        if (!write_point(cm)) H2A_FAIL(ctx, H2A_ERR_INVALID, "create_proof: the random polynomial commits to the identity");                                                           // :419-421
    }
    hh::Fr y = tr.squeeze();                                                       // :423

    // coefficient and extended forms of everything the quotient touches: queued on the transform lane as each group of
    // columns became final (transform_columns above); this phase is what is left of them when the commitments are done
    H2A_TRY(join_transforms(ctx, c));
    steps.mark("ifft + coset fft of committed columns");
    hh::fr_store(p->qargs.beta, beta); hh::fr_store(p->qargs.gamma, gamma); hh::fr_store(p->qargs.theta, theta); hh::fr_store(p->qargs.y, y);
    {   // weights of the quotient's expressions, in the kernel's order: expression k of K carries y^(K-1-k)
        const size_t K = s.gates.size() + (s.n_chunks ? 2 + (s.n_chunks - 1) + s.n_chunks : 0) + 5 * s.lookups.size();
        if (K > (size_t)MAX_EXPR) H2A_FAIL(ctx, H2A_ERR_INVALID, "create_proof: %zu quotient expressions exceed %d", K, MAX_EXPR);
        hh::Fr wgt = hh::fr_one();
        for (size_t e = K; e-- > 0;) { hh::fr_store(p->qargs.ypow[e], wgt); wgt = wgt * y; }
    }
    // several GPUs: rank r evaluates rows [r m / W, (r + 1) m / W) and the slices are allgathered in place
    const bool q_rows = p->dist_slice != 0;
    const uint32_t q_cnt = q_rows ? p->dist_slice : m;
    p->qargs.row_lo = q_rows ? q_cnt * (uint32_t)c->dist_rank : 0u;
    p->qargs.row_hi = p->qargs.row_lo + q_cnt;
    H2A_CUDA(ctx, cudaMemcpyAsync(p->d_qargs, &p->qargs, sizeof(QuotientArgs), cudaMemcpyHostToDevice, st));
    {
        static const int minb = getenv("H2A_QUOTIENT_MINB") ? atoi(getenv("H2A_QUOTIENT_MINB")) : 3;
        if (minb >= 5) LAUNCH1D(dev::quotient_kernel<5>, q_cnt, 128, p->d_qargs);
        else if (minb == 4) LAUNCH1D(dev::quotient_kernel<4>, q_cnt, 128, p->d_qargs);
        else LAUNCH1D(dev::quotient_kernel<3>, q_cnt, 128, p->d_qargs);
    }
    if (q_rows) H2A_TRY(h2a_comm_allgather_on(ctx, 0, p->h_ext + 32ull * p->qargs.row_lo, p->h_ext, 32ull * q_cnt, st));
    {   // extended_to_coeff, then h is cut into quotient_poly_degree pieces of n coefficients
        uint8_t w[32];
        hh::fr_store(w, hh::fr_root_of_unity((int)s.ext_k));
        H2A_TRY(h2a_ntt_run(ctx, p->h_ext, m, p->h_ext, p->h_coef, s.ext_k, w, 1, p->coset));
    }
    steps.mark("quotient evaluation + extended_to_coeff");
    {                                                                              // :427-434, src/vanishing.rs:77-106
        std::vector<const uint8_t*> cols;
        for (uint32_t i = 0; i < s.qdeg; i++) cols.push_back(p->h_coef + 32ull * n * i);
        std::vector<hh::PointA> cms;
        H2A_TRY(commit_batch(ctx, c, p->g, cols, n, cms));
        H2A_TRY(write_points(cms, "quotient piece"));
    }
    steps.mark("h commitments");
    hh::Fr x = tr.squeeze();                                                       // :436

    // ---- evaluations: all requested first, fetched with one copy
    struct Ev { const uint8_t* coef; int32_t rot; };
    std::vector<Ev> evs;
    for (auto& q : s.iq) evs.push_back({p->instance[q.col].coef, q.rot});
    for (auto& q : s.aq) evs.push_back({p->advice[q.col].coef, q.rot});
    for (auto& q : s.fq) evs.push_back({p->fixed[q.col].coef, q.rot});
    evs.push_back({p->random_coef, 0});
    for (auto& q : p->sigma) evs.push_back({q.coef, 0});
    for (uint32_t i = 0; i < s.n_chunks; i++) {
        evs.push_back({p->pz[i].coef, 0});
        evs.push_back({p->pz[i].coef, 1});
        if (i + 1 < s.n_chunks) evs.push_back({p->pz[i].coef, s.last_rot});
    }
    for (auto& l : p->lk) {
        evs.push_back({l.z.coef, 0}); evs.push_back({l.z.coef, 1}); evs.push_back({l.pa.coef, 0}); evs.push_back({l.pa.coef, -1}); evs.push_back({l.ps.coef, 0});
    }
    if (evs.size() > (size_t)MAX_EVALS) H2A_FAIL(ctx, H2A_ERR_INVALID, "create_proof: too many evaluations");
    // the evaluation points x * omega^rot live AFTER the evaluation results, so any number of distinct rotations fits as long
    // as the 1024-slot scalar buffer does (a fixed block in front of the results would run into them at the 9th rotation)
    const int S_POINT0 = S_EVAL0 + (int)evs.size();
    std::map<int32_t, int> point_slot;                                            // rotation -> slot of x * omega^rot
    auto point_of = [&](int32_t rot) -> int {
        auto it = point_slot.find(rot);
        if (it != point_slot.end()) return it->second;
        int sl = S_POINT0 + (int)point_slot.size();
        point_slot[rot] = sl;
        return sl;
    };
    for (auto& e : evs) point_of(e.rot);
    point_of(0);
    if ((size_t)S_POINT0 + point_slot.size() > 1024) H2A_FAIL(ctx, H2A_ERR_INVALID, "create_proof: %zu evaluations and %zu rotations exceed the scalar buffer", evs.size(), point_slot.size());
    for (auto& kv : point_slot) H2A_TRY(upload_fr(ctx, slot(kv.second), rotate_point(s, x, kv.first)));
    {
        // several GPUs: evaluation i is computed by rank i % world (every rank holds every coefficient form); the 32-byte results
        // are exchanged below
        std::vector<dev::EvalReq> reqs;
        for (size_t i = 0; i < evs.size(); i++)
            if (!dist || (int)(i % (size_t)dist_w) == dist_me) reqs.push_back(dev::EvalReq{evs[i].coef, slot(point_slot[evs[i].rot]), slot(S_EVAL0 + (int)i)});
        H2A_CUDA(ctx, cudaMemcpyAsync(p->d_reqs, reqs.data(), sizeof(dev::EvalReq) * reqs.size(), cudaMemcpyHostToDevice, st));
        H2A_CUDA(ctx, cudaStreamSynchronize(st));   // reqs is a stack-owned vector
        const uint32_t nch = (n + dev::HORNER_L - 1) / dev::HORNER_L;
        if (!reqs.empty()) {
            dim3 grid((nch + 127) / 128, (unsigned)reqs.size());
            dev::chunk_values_multi_kernel<<<grid, 128, 0, st>>>(p->d_reqs, n, nch, p->eval_chunks);
            H2A_LAUNCH_CHECK(ctx);
            dev::fold_chunks_multi_kernel<<<(unsigned)reqs.size(), 256, 0, st>>>(p->d_reqs, nch, p->eval_chunks);
            H2A_LAUNCH_CHECK(ctx);
        }
    }
    std::vector<uint8_t> evb(32 * evs.size());
    H2A_CUDA(ctx, cudaMemcpyAsync(evb.data(), slot(S_EVAL0), evb.size(), cudaMemcpyDeviceToHost, st));
    H2A_CUDA(ctx, cudaStreamSynchronize(st));
    if (dist) {   // slot i of rank i % world is the one that was computed
        std::vector<uint8_t> all(evb.size() * (size_t)dist_w);
        H2A_TRY(h2a_comm_allgather(ctx, evb.data(), all.data(), evb.size()));
        for (size_t i = 0; i < evs.size(); i++) memcpy(evb.data() + 32 * i, all.data() + evb.size() * (i % (size_t)dist_w) + 32 * i, 32);
    }
    std::vector<hh::Fr> evals(evs.size());
    for (size_t i = 0; i < evs.size(); i++) { evals[i] = hh::fr_load(evb.data() + 32 * i); write_scalar(evals[i]); }   // :438-510
    steps.mark("evaluations");
    hh::Fr v = tr.squeeze();                                                       // :718
    (void)tr.squeeze();                                                            // :719 u (not used by the prover)
    H2A_TRY(upload_fr(ctx, slot(S_V), v));

    // ---- multi-open: queries in the verifier's order (:654-715), grouped by rotation (src/multiopen.rs:19-45)
    // H = sum_i (x^n)^i h_i as a polynomial of degree < n (Horner from the top piece)
    hh::Fr xn = x;
    for (uint32_t i = 0; i < s.k; i++) xn = hh::sqr(xn);
    H2A_TRY(upload_fr(ctx, slot(S_Y), xn));
    uint8_t* h_poly = p->tmp_n[0];
    for (int i = (int)s.qdeg - 1; i >= 0; i--)
        LAUNCH1D(dev::axpy_kernel, n, 256, h_poly, p->h_coef + 32ull * n * i, slot(S_Y), n, i == (int)s.qdeg - 1);
    struct MQ { const uint8_t* coef; int32_t rot; };
    std::vector<MQ> mq;
    for (auto& q : s.iq) mq.push_back({p->instance[q.col].coef, q.rot});
    for (auto& q : s.aq) mq.push_back({p->advice[q.col].coef, q.rot});
    for (uint32_t i = 0; i < s.n_chunks; i++) { mq.push_back({p->pz[i].coef, 0}); mq.push_back({p->pz[i].coef, 1}); }
    for (int i = (int)s.n_chunks - 2; i >= 0; i--) mq.push_back({p->pz[i].coef, s.last_rot});
    for (auto& l : p->lk) { mq.push_back({l.z.coef, 0}); mq.push_back({l.pa.coef, 0}); mq.push_back({l.ps.coef, 0}); mq.push_back({l.pa.coef, -1}); mq.push_back({l.z.coef, 1}); }
    for (auto& q : s.fq) mq.push_back({p->fixed[q.col].coef, q.rot});
    for (auto& q : p->sigma) mq.push_back({q.coef, 0});
    mq.push_back({h_poly, 0});
    mq.push_back({p->random_coef, 0});
    std::map<int32_t, std::vector<size_t>> sets;
    for (size_t i = 0; i < mq.size(); i++) sets[mq[i].rot].push_back(i);
    const uint32_t nchunks = (n + dev::HORNER_L - 1) / dev::HORNER_L;
    {                                                                              // src/multiopen.rs:344-395 (prover mirror)
        std::vector<const uint8_t*> wcols;
        std::vector<hh::PointA> cms;
        const size_t slots = m / n;                                                // quotients parked in the (now free) h_ext
        size_t used = 0;
        for (auto& kv : sets) {
            uint8_t* batch = p->tmp_n[1];
            if (!point_slot.count(kv.first)) H2A_FAIL(ctx, H2A_ERR_INVALID, "create_proof: rotation without evaluation point");
            const uint8_t* d_z = slot(point_slot[kv.first]);
            if (used == slots) {                                                   // more rotation sets than slots: flush
                std::vector<hh::PointA> part;
                H2A_TRY(commit_batch(ctx, c, p->g, wcols, n, part));
                cms.insert(cms.end(), part.begin(), part.end());
                wcols.clear();
                used = 0;
            }
            // several GPUs: the rotation set in position j of a batch of commitments belongs to rank j % world (the owner
            // commit_batch gives column j): only that rank folds the set's polynomials and divides
            const bool mine = !dist || (int)(used % (size_t)dist_w) == dist_me;
            uint8_t* q = p->h_ext + 32ull * n * used++;
            wcols.push_back(q);
            if (!mine) continue;
            bool first = true;
            for (size_t qi : kv.second) { LAUNCH1D(dev::axpy_kernel, n, 256, batch, mq[qi].coef, slot(S_V), n, first ? 1 : 0); first = false; }
            LAUNCH1D(dev::chunk_values_kernel, nchunks, 128, batch, n, d_z, p->chunks);
            dev::kate_carry_kernel<<<1, 128, 0, st>>>(p->chunks, nchunks, d_z, p->carries);
            H2A_LAUNCH_CHECK(ctx);
            LAUNCH1D(dev::kate_quotient_kernel, nchunks, 128, batch, n, d_z, p->carries, q);
        }
        std::vector<hh::PointA> part;
        H2A_TRY(commit_batch(ctx, c, p->g, wcols, n, part));
        cms.insert(cms.end(), part.begin(), part.end());
        for (auto& cm : cms) {                                                     // src/multiopen.rs:392; not absorbed (:202-218)
            if (hh::is_identity(cm)) H2A_FAIL(ctx, H2A_ERR_INVALID, "create_proof: an opening witness W is the identity");
            compress_point(cm, proof_out + pos);
            pos += 32;
        }
    }
    steps.mark("multiopen witnesses");
    *proof_len = pos;

    H2A_CUDA(ctx, cudaStreamSynchronize(st));
    p->phase_ms.clear();
    ctx->prove_phase_names.clear();
    for (size_t i = 1; i < steps.ev.size(); i++) {
        float ms = 0;
        cudaEventElapsedTime(&ms, steps.ev[i - 1], steps.ev[i]);
        p->phase_ms.push_back(ms);
        ctx->prove_phase_names.push_back(steps.names[i]);
    }
    return H2A_OK;   // the guard releases the events
}

// KZG setup with a caller-supplied secret: g[i] = [s^i] G, g_lagrange[i] = [L_i(s)] G, L_i(s) = omega^i (s^n - 1) / (n (s - omega^i)).
int h2a_kzg_setup(h2a_ctx* ctx, uint32_t k, const uint8_t s_[32], h2a_bases** out_g, h2a_bases** out_g_lagrange) {
    H2A_DEVICE(ctx);
    if (!ctx || !s_ || !out_g || !out_g_lagrange) return H2A_ERR_INVALID;
    if (k < 1 || k > 26) H2A_FAIL(ctx, H2A_ERR_INVALID, "kzg_setup: k=%u not in 1..26", k);
    const uint32_t n = 1u << k;
    cudaStream_t st = ctx->stream;
    // table of 2^j * G
    std::vector<uint8_t> tab(254 * 64);
    {
        hh::PointX cur = hh::px_from_affine(hh::PointA{hh::fq_one(), hh::Fq{hh::el_from_u64(2, hh::MOD_Q)}});
        for (int j = 0; j < 254; j++) {
            hh::affine_store(tab.data() + 64 * j, hh::px_to_affine(cur));
            cur = hh::px_dbl(cur);
        }
    }
    uint8_t *d_tab = nullptr, *d_sc = nullptr, *d_w = nullptr, *d_den = nullptr, *d_small = nullptr, *d_g = nullptr, *d_gl = nullptr;
    auto cleanup = [&]() { for (uint8_t* q : {d_tab, d_sc, d_w, d_den, d_small}) if (q) cudaFree(q); };
#define SETUP_CUDA(call)                                                                    \
    do {                                                                                    \
        cudaError_t _e = (call);                                                            \
        if (_e != cudaSuccess) {                                                            \
            cleanup();                                                                      \
            if (d_g) cudaFree(d_g);                                                         \
            if (d_gl) cudaFree(d_gl);                                                       \
            H2A_FAIL(ctx, H2A_ERR_CUDA, "kzg_setup: %s -> %s", #call, cudaGetErrorString(_e)); \
        }                                                                                   \
    } while (0)
    SETUP_CUDA(cudaMalloc(&d_tab, tab.size()));
    SETUP_CUDA(cudaMalloc(&d_sc, 32ull * n));
    SETUP_CUDA(cudaMalloc(&d_w, 32ull * n));
    SETUP_CUDA(cudaMalloc(&d_den, 32ull * n));
    SETUP_CUDA(cudaMalloc(&d_small, 256));
    SETUP_CUDA(cudaMalloc(&d_g, 64ull * n));
    SETUP_CUDA(cudaMalloc(&d_gl, 64ull * n));
    SETUP_CUDA(cudaMemcpyAsync(d_tab, tab.data(), tab.size(), cudaMemcpyHostToDevice, st));
    hh::Fr s = hh::fr_load(s_), omega = hh::fr_root_of_unity((int)k), sn = s;
    for (uint32_t i = 0; i < k; i++) sn = hh::sqr(sn);
    uint8_t small[96], oneb[32], wb[32];
    hh::fr_store(small, s);
    hh::fr_store(small + 32, hh::fr_from_u64(n));
    hh::fr_store(small + 64, sn - hh::fr_one());
    hh::fr_store(oneb, hh::fr_one());
    hh::fr_store(wb, omega);
    SETUP_CUDA(cudaMemcpyAsync(d_small, small, 96, cudaMemcpyHostToDevice, st));
    SETUP_CUDA(cudaStreamSynchronize(st));
    int rc = h2a_pow_vector(ctx, s_, oneb, n, d_sc);                       // s^i
    if (rc == H2A_OK) rc = h2a_pow_vector(ctx, wb, oneb, n, d_w);          // omega^i
    if (rc != H2A_OK) { cleanup(); cudaFree(d_g); cudaFree(d_gl); return rc; }
    const unsigned blocks = (n + 127) / 128;
    dev::fixed_base_mul_kernel<<<blocks, 128, 0, st>>>(d_tab, d_sc, n, d_g);
    ctx->launches++;
    dev::lagrange_den_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_w, d_small, d_small + 32, n, d_den);
    dev::batch_inverse_kernel<<<((n + dev::INV_CHUNK - 1) / dev::INV_CHUNK + 127) / 128, 128, 0, st>>>(d_den, n);
    dev::lagrange_num_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_w, d_small + 64, d_den, n, d_sc);
    dev::fixed_base_mul_kernel<<<blocks, 128, 0, st>>>(d_tab, d_sc, n, d_gl);
    ctx->launches += 4;
    SETUP_CUDA(cudaGetLastError());
    SETUP_CUDA(cudaStreamSynchronize(st));
#undef SETUP_CUDA
    cleanup();
    h2a_bases *bg = new h2a_bases(), *bl = new h2a_bases();
    bg->d = d_g; bg->n = n; bg->owned = true;
    bl->d = d_gl; bl->n = n; bl->owned = true;
    *out_g = bg;
    *out_g_lagrange = bl;
    return H2A_OK;
}

// Per-phase times (ms) of the last h2a_create_proof on this circuit; returns the count written.
int h2a_prove_phase_ms(h2a_ctx* ctx, const h2a_circuit* c, float* ms, int cap) {
    if (!ctx || !c || !c->prover || !ms) return H2A_ERR_INVALID;
    int k = std::min<int>(cap, (int)c->prover->phase_ms.size());
    for (int i = 0; i < k; i++) ms[i] = c->prover->phase_ms[i];
    return k;
}
const char* h2a_prove_phase_name(const h2a_ctx* ctx, int i) {
    if (!ctx || i < 0 || i >= (int)ctx->prove_phase_names.size()) return "";
    return ctx->prove_phase_names[i];
}

}  // extern "C"
