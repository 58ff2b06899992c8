// ntt.cu — Fr number-theoretic transform for sm_100a (BN254 scalar field, 2-adicity 28).
//
// Computes what halo2 `arithmetic::best_fft(a, omega, log_n)` computes — natural order in and out,
// a[i] <- sum_j a[j] omega^(ij) — plus the `EvaluationDomain` variants built on it (inverse with 1/n,
// coset shift, zero-padded extension).  The reference touches the domain at src/verifier.rs:252,431;
// the transforms themselves run inside `create_proof` (examples/simple-example.rs:606-613, :702-709).
//
// Decomposition: n = n_1 * n_2 (* n_3), n_p = 2^s_p <= 1024.  Pass p transforms digit p of the index
// inside shared memory (radix-2 DIF butterflies, twiddles from a shared-memory table), multiplies by
// the inter-pass twist omega_m^(i_p * inner) read from a resident table of omega powers, and stores in
// place; the last pass writes the digit-reversed (= natural) order to a second buffer.  Zero padding,
// the coset pre-scale g^j and the inverse post-scale g^-j / n are fused into the first load and the
// last store.  Each pass reads and writes every element once: 64 B of HBM traffic per element per pass.
#include <cstring>
#include <string>

#include "ctx.hpp"
#include "field.cuh"
#include "host_bn254.hpp"

using namespace h2a;

struct NttTables {
    uint32_t log_n = 0;
    uint8_t omega[32];
    DevBuf tw;  // omega^i, i < n
    // cached power tables for coset scaling: key = 32-byte base || log_n || mode
    std::map<std::string, DevBuf> pow_tables;
};

namespace {

constexpr int NTT_THREADS = 256;   // upper bound; a launch uses one thread per register group of the tile
// butterfly levels a register group takes at once (2: radix-4 groups, 4 elements per thread; 3: radix-8, 8 elements)
#ifndef NTT_NB
#define NTT_NB 2
#endif
#ifndef NTT_MINB
#define NTT_MINB 3
#endif
constexpr int LOG_PW_LO = 10;

struct PassArgs {
    const uint8_t* src;
    uint8_t* dst;
    uint32_t log_n;     // transform size
    uint32_t s;         // this pass transforms 2^s-point sub-transforms
    uint32_t log_r;     // log2 of the remaining size after this pass (0 for the last pass)
    uint32_t log_tile;  // sub-transforms per block
    uint32_t n_in;      // elements present in src; indices beyond read as zero
    const uint8_t* tw;  // omega^i table
    uint32_t tw_shift;  // table holds omega_T^i with T = n << tw_shift
    int inverse;
    int pre_mode;   // 0 none, 1 multiply input j by pw(j)
    int post_mode;  // 0 none, 1 multiply output j by constant pw_lo[0], 2 by pw(j)
    const uint8_t* pw_lo;  // base^i (times a constant for post), i < 2^LOG_PW_LO
    const uint8_t* pw_hi;  // base^(i << LOG_PW_LO)
    uint32_t s1, s2;       // radix bits of passes 1 and 2 (last pass only; 0 when absent)
    int bulk;              // last pass: rows staged by TMA bulk copies (see ntt_last_pass_kernel)
};

// shared-memory element storage: two planes of uint4 so that consecutive elements are 16 B apart.  SWZ: the low three
// bits of the index are XORed with the next three, so that the strided accesses of the register groups below (8 lanes of a
// quarter-warp reading elements 4 or 8 apart) still fall into eight different 16-byte bank groups.
template <bool SWZ>
struct SmemViewT {
    uint4* lo;
    uint4* hi;
    __device__ __forceinline__ static uint32_t at(uint32_t i) { return SWZ ? i ^ ((i >> 3) & 7u) : i; }
    __device__ __forceinline__ Fr get(uint32_t i_) const {
        const uint32_t i = at(i_);
        Fr r;
        uint4 a = lo[i], b = hi[i];
        r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
        r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
        return r;
    }
    __device__ __forceinline__ void put(uint32_t i_, const Fr& v) const {
        const uint32_t i = at(i_);
        lo[i] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
        hi[i] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
    }
};
typedef SmemViewT<true> SmemView;      // data tile
typedef SmemViewT<false> TwView;       // twiddle table

__device__ __forceinline__ Fr tw_lookup(const PassArgs& a, uint32_t e) {  // omega^(+-e), e < n
    uint32_t n_mask = (1u << a.log_n) - 1u;
    uint32_t idx = a.inverse ? ((0u - e) & n_mask) : e;
    return Fr::load(a.tw + 32ull * ((size_t)idx << a.tw_shift));
}
__device__ __forceinline__ Fr pw_lookup(const PassArgs& a, uint32_t j) {
    Fr lo = Fr::load(a.pw_lo + 32ull * (j & ((1u << LOG_PW_LO) - 1u)));
    uint32_t h = j >> LOG_PW_LO;
    if (h == 0) return lo;
    return lo * Fr::load(a.pw_hi + 32ull * h);
}
__device__ __forceinline__ Fr load_input(const PassArgs& a, uint32_t gidx) {
    if (gidx >= a.n_in) return Fr::zero();
    Fr v = Fr::load(a.src + 32ull * gidx);
    if (a.pre_mode) v = v * pw_lookup(a, gidx);
    return v;
}

// DIF stages lh = b + NB - 1 .. b of `rows` sub-transforms of 2^s points held in shared memory, NB stages at a time IN
// REGISTERS: a thread takes the 2^NB elements of one row whose digits differ in bits [b, b + NB), runs the NB butterfly levels
// on them without touching shared memory, and puts them back — one shared-memory round trip and one barrier per NB levels
// instead of per level, and the index arithmetic of a work item is shared by NB * 2^(NB-1) butterflies.
// Element (row t, digit d) lives at d*dstride + t*tstride.  Output digit i ends at position bitrev_s(i).
// `stage` (optional): the group's inputs are read from there — a linear array of 32-byte elements with the same logical indexing,
// filled by TMA bulk copies (ntt_last_pass_kernel) — instead of from the planes; the results always go to the planes.
template <int NB>
__device__ __forceinline__ void reg_group(const SmemView& sm, const TwView& ltw, uint32_t s, uint32_t b, uint32_t log_rows,
                                          uint32_t dstride, uint32_t tstride, bool rows_fastest, const uint8_t* stage = nullptr) {
    constexpr uint32_t E = 1u << NB;
    const uint32_t log_q = s - NB;                       // digits left to the work item index
    const uint32_t items = 1u << (log_q + log_rows);
    for (uint32_t w = threadIdx.x; w < items; w += blockDim.x) {
        uint32_t t, q;
        if (rows_fastest) { t = w & ((1u << log_rows) - 1u); q = w >> log_rows; }
        else { q = w & ((1u << log_q) - 1u); t = w >> log_q; }
        const uint32_t q_lo = q & ((1u << b) - 1u);
        const uint32_t d_base = ((q >> b) << (b + NB)) | q_lo;
        const uint32_t i_base = d_base * dstride + t * tstride, i_step = dstride << b;
        Fr v[E];
#pragma unroll
        for (uint32_t j = 0; j < E; j++) v[j] = stage ? Fr::load(stage + 32u * (i_base + j * i_step)) : sm.get(i_base + j * i_step);
#pragma unroll
        for (int st = NB - 1; st >= 0; st--) {
            const uint32_t half = 1u << st, lh = b + (uint32_t)st;
#pragma unroll
            for (uint32_t j0 = 0; j0 < E; j0++) {
                if (j0 & half) continue;
                const uint32_t j1 = j0 | half;
                const uint32_t pos = q_lo | ((j0 & (half - 1u)) << b);     // digit of the pair below bit lh
                const Fr u = v[j0], x = v[j1];
                v[j0] = u + x;
                Fr d = u - x;
                if (pos) d = d * ltw.get(pos << (s - 1u - lh));
                v[j1] = d;
            }
        }
#pragma unroll
        for (uint32_t j = 0; j < E; j++) sm.put(i_base + j * i_step, v[j]);
    }
    __syncthreads();
}
// all s levels, NTT_NB at a time from the top (the remainder comes last)
__device__ __forceinline__ void smem_dif(const SmemView& sm, const TwView& ltw, uint32_t s, uint32_t log_rows,
                                         uint32_t dstride, uint32_t tstride, bool rows_fastest, const uint8_t* stage = nullptr) {
    uint32_t top = s;                                    // levels [0, top) are still to do
    while (top >= NTT_NB) { top -= NTT_NB; reg_group<NTT_NB>(sm, ltw, s, top, log_rows, dstride, tstride, rows_fastest, stage); stage = nullptr; }
#if NTT_NB == 3
    if (top == 2) { reg_group<2>(sm, ltw, s, 0, log_rows, dstride, tstride, rows_fastest, stage); stage = nullptr; }
#endif
    if (top == 1) reg_group<1>(sm, ltw, s, 0, log_rows, dstride, tstride, rows_fastest, stage);
}

__device__ __forceinline__ void load_local_twiddles(const PassArgs& a, const TwView& ltw) {
    // ltw[e] = omega_{n_p}^e = omega^(e * n / n_p), e < n_p/2
    const uint32_t half_n = 1u << (a.s - 1);
    for (uint32_t e = threadIdx.x; e < half_n; e += blockDim.x) ltw.put(e, tw_lookup(a, e << (a.log_n - a.s)));
}

// Passes 1..P-1: the transformed digit has stride R = 2^log_r; a block takes `tile` consecutive inner
// positions.  Shared layout: d * tile + t.
__global__ void __launch_bounds__(NTT_THREADS, NTT_MINB) ntt_strided_pass_kernel(PassArgs a) {
    extern __shared__ uint4 smem_raw[];
    const uint32_t np = 1u << a.s, tile = 1u << a.log_tile, elems = np << a.log_tile;
    SmemView sm{smem_raw, smem_raw + elems};
    TwView ltw{smem_raw + 2 * elems, smem_raw + 2 * elems + (np >> 1)};
    const uint32_t tiles_per_outer = 1u << (a.log_r - a.log_tile);
    const uint32_t outer = blockIdx.x / tiles_per_outer;
    const uint32_t inner0 = (blockIdx.x % tiles_per_outer) << a.log_tile;
    const size_t base = ((size_t)outer << (a.s + a.log_r)) + inner0;

    load_local_twiddles(a, ltw);
    for (uint32_t x = threadIdx.x; x < elems; x += blockDim.x) {
        uint32_t d = x >> a.log_tile, t = x & (tile - 1u);
        sm.put(x, load_input(a, (uint32_t)(base + ((size_t)d << a.log_r) + t)));
    }
    __syncthreads();
    smem_dif(sm, ltw, a.s, a.log_tile, tile, 1, true);
    // twist by omega_m^(i_p * inner), m = n_p * R, omega_m = omega^(n/m)
    const uint32_t log_m = a.s + a.log_r;
    for (uint32_t x = threadIdx.x; x < elems; x += blockDim.x) {
        uint32_t dpos = x >> a.log_tile, t = x & (tile - 1u);
        uint32_t ip = __brev(dpos) >> (32 - a.s);
        uint32_t inner = inner0 + t;
        Fr v = sm.get(x);
        uint32_t e = (ip * inner) << (a.log_n - log_m);
        if (e) v = v * tw_lookup(a, e);
        v.store(a.dst + 32ull * (base + ((size_t)ip << a.log_r) + t));
    }
}

// Last pass: contiguous 2^s-point sub-transforms; a block takes `tile` of them that differ in digit 1
// so that the digit-reversed stores form runs of `tile` consecutive outputs.
// Shared layout: t * n_p + d (swizzled, see SmemViewT).
__global__ void __launch_bounds__(NTT_THREADS, NTT_MINB) ntt_last_pass_kernel(PassArgs a) {
    extern __shared__ uint4 smem_raw[];
    const uint32_t np = 1u << a.s, tile = 1u << a.log_tile, row = np, elems = row << a.log_tile;
    SmemView sm{smem_raw, smem_raw + elems};
    TwView ltw{smem_raw + 2 * elems, smem_raw + 2 * elems + (np >> 1)};
    const uint32_t log_outer = a.log_n - a.s;             // number of sub-transforms = 2^log_outer
    const uint32_t log_rest = log_outer - a.s1;           // digits between digit 1 and the last one
    const uint32_t i_rest = blockIdx.x & ((1u << log_rest) - 1u);
    const uint32_t i1_0 = (blockIdx.x >> log_rest) << a.log_tile;

    // The rows of this pass are contiguous in global memory (n_p elements = 32 n_p bytes each): when nothing has to happen to
    // an element on its way in (no coset pre-scale, no zero padding) one elected thread asks the TMA unit for them — one 1-D bulk
    // copy per row into a linear staging area, completion counted in bytes on an mbarrier — while the block builds its twiddle
    // table; the first register group then reads its inputs from the staging area.  Otherwise every thread loads its elements.
    uint8_t* stage = nullptr;
    if (a.bulk) {
        stage = (uint8_t*)(smem_raw + 2 * elems + 2 * (np >> 1) + 1);            // after the twiddle planes; 16-byte aligned
        uint64_t* bar = (uint64_t*)(smem_raw + 2 * elems + 2 * (np >> 1));
        const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(bar), row_bytes = 32u << a.s;
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(row_bytes << a.log_tile) : "memory");
            for (uint32_t t = 0; t < tile; t++) {
                const uint32_t outer = ((i1_0 + t) << log_rest) + i_rest;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 (uint32_t)__cvta_generic_to_shared(stage + (size_t)t * row_bytes)),
                             "l"(a.src + 32ull * ((size_t)outer << a.s)), "r"(row_bytes), "r"(bar_s)
                             : "memory");
            }
        }
        load_local_twiddles(a, ltw);
        __syncthreads();                                                         // twiddles written, barrier initialised
        uint32_t done;
        do {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar_s) : "memory");
        } while (!done);
    } else {
        load_local_twiddles(a, ltw);
        for (uint32_t x = threadIdx.x; x < (np << a.log_tile); x += blockDim.x) {
            uint32_t t = x >> a.s, d = x & (np - 1u);
            uint32_t outer = ((i1_0 + t) << log_rest) + i_rest;
            sm.put(t * row + d, load_input(a, (outer << a.s) + d));
        }
        __syncthreads();
    }
    smem_dif(sm, ltw, a.s, a.log_tile, 1, row, false, stage);
    for (uint32_t x = threadIdx.x; x < (np << a.log_tile); x += blockDim.x) {
        uint32_t t = x & (tile - 1u), dpos = x >> a.log_tile;
        uint32_t ip = a.s ? (__brev(dpos) >> (32 - a.s)) : 0u;
        // output index: i_1 + n_1 * i_2 + (n / n_P) * i_P   (i_rest is digit 2 when there are 3 passes)
        uint32_t out = (i1_0 + t) + (i_rest << a.s1) + (ip << log_outer);
        Fr v = sm.get(t * row + dpos);
        if (a.post_mode == 1) v = v * Fr::load(a.pw_lo);
        else if (a.post_mode == 2) v = v * pw_lookup(a, out);
        v.store(a.dst + 32ull * out);
    }
}

// tables of powers: lo[i] = c * base^i (i < 2^LOG_PW_LO), hi[i] = base^(i << LOG_PW_LO)
__global__ void pow_table_kernel(const uint8_t* base_c /* base || c */, uint32_t n_lo, uint32_t n_hi, uint8_t* lo,
                                 uint8_t* hi) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_lo + n_hi) return;
    Fr base = Fr::load(base_c);
    uint32_t e[1];
    if (i < n_lo) {
        e[0] = i;
        Fr r = base.pow_limbs(e, 32 - __clz(i | 1)) * Fr::load(base_c + 32);
        r.store(lo + 32ull * i);
    } else {
        uint32_t j = i - n_lo;
        e[0] = j << LOG_PW_LO;
        Fr r = base.pow_limbs(e, 32 - __clz(e[0] | 1));
        r.store(hi + 32ull * j);
    }
}
__global__ void expand_pow_kernel(const uint8_t* lo, const uint8_t* hi, uint32_t n, uint8_t* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr v = Fr::load(lo + 32ull * (i & ((1u << LOG_PW_LO) - 1u)));
    uint32_t h = i >> LOG_PW_LO;
    if (h) v = v * Fr::load(hi + 32ull * h);
    v.store(out + 32ull * i);
}

// Builds lo/hi power tables for `base` (scaled by `c`) into one buffer: lo at 0, hi at 32 << LOG_PW_LO.
int build_pow_tables(h2a_ctx* ctx, const h2a_host::Fr& base, const h2a_host::Fr& c, uint32_t n, DevBuf& out) {
    const uint32_t n_lo = 1u << LOG_PW_LO;
    const uint32_t n_hi = std::max(1u, (n + n_lo - 1) >> LOG_PW_LO);
    H2A_TRY(h2a_reserve(ctx, out, 32ull * (n_lo + n_hi) + 64));
    uint8_t* stage = (uint8_t*)out.p + 32ull * (n_lo + n_hi);
    uint8_t host[64];
    h2a_host::fr_store(host, base);
    h2a_host::fr_store(host + 32, c);
    H2A_CUDA(ctx, cudaMemcpyAsync(stage, host, 64, cudaMemcpyHostToDevice, ctx->stream));
    H2A_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `host` is a stack buffer
    pow_table_kernel<<<(n_lo + n_hi + 127) / 128, 128, 0, ctx->stream>>>(stage, n_lo, n_hi, (uint8_t*)out.p,
                                                                         (uint8_t*)out.p + 32ull * n_lo);
    H2A_LAUNCH_CHECK(ctx);
    return H2A_OK;
}

int get_tables(h2a_ctx* ctx, uint32_t log_n, const uint8_t omega[32], NttTables** out) {
    // reuse any cached table whose omega generates ours: omega == tab.omega^(2^(tab.log_n - log_n))
    for (auto& kv : ctx->ntt_tables) {
        NttTables* t = kv.second;
        if (t->log_n < log_n) continue;
        h2a_host::Fr w = h2a_host::fr_load(t->omega);
        for (uint32_t i = log_n; i < t->log_n; i++) w = h2a_host::sqr(w);
        uint8_t wb[32];
        h2a_host::fr_store(wb, w);
        if (memcmp(wb, omega, 32) == 0) { *out = t; return H2A_OK; }
    }
    NttTables* t = new NttTables();
    t->log_n = log_n;
    memcpy(t->omega, omega, 32);
    const uint32_t n = 1u << log_n;
    DevBuf pw;
    int rc = build_pow_tables(ctx, h2a_host::fr_load(omega), h2a_host::fr_one(), n, pw);
    if (rc == H2A_OK) rc = h2a_reserve(ctx, t->tw, 32ull * n);
    if (rc != H2A_OK) { if (pw.p) cudaFree(pw.p); delete t; return rc; }
    expand_pow_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>((uint8_t*)pw.p, (uint8_t*)pw.p + (32ull << LOG_PW_LO), n,
                                                                (uint8_t*)t->tw.p);
    ctx->launches++;
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(pw.p);
    if (e != cudaSuccess) { cudaFree(t->tw.p); delete t; H2A_FAIL(ctx, H2A_ERR_CUDA, "ntt tables: %s", cudaGetErrorString(e)); }
    // key: unique per (log_n, omega) — several omegas may share a log_n
    uint32_t key = log_n;
    while (ctx->ntt_tables.count(key)) key += 64;
    ctx->ntt_tables[key] = t;
    *out = t;
    return H2A_OK;
}

int get_pow_tables(h2a_ctx* ctx, NttTables* t, const h2a_host::Fr& base, const h2a_host::Fr& c, uint32_t n, DevBuf** out) {
    std::string key((const char*)base.e.v, 32);
    key.append((const char*)c.e.v, 32);
    key.append((const char*)&n, 4);
    auto it = t->pow_tables.find(key);
    if (it != t->pow_tables.end()) { *out = &it->second; return H2A_OK; }
    if (t->pow_tables.size() > 16) {  // bounded cache
        H2A_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (auto& kv : t->pow_tables) cudaFree(kv.second.p);
        t->pow_tables.clear();
    }
    DevBuf b;
    H2A_TRY(build_pow_tables(ctx, base, c, n, b));
    t->pow_tables[key] = b;
    *out = &t->pow_tables[key];
    return H2A_OK;
}

void split_bits(uint32_t k, uint32_t s[3], int* passes) {
    if (k <= 10) { s[0] = k; s[1] = s[2] = 0; *passes = 1; }
    else if (k <= 20) { s[0] = (k + 1) / 2; s[1] = k / 2; s[2] = 0; *passes = 2; }
    else { s[0] = (k + 2) / 3; s[1] = (k + 1) / 3; s[2] = k / 3; *passes = 3; }
}

// one thread per register group of the tile, whole warps, at most NTT_THREADS
unsigned threads_for(uint32_t log_elems) { return (unsigned)std::min<uint32_t>(NTT_THREADS, std::max<uint32_t>(32u, (1u << log_elems) >> NTT_NB)); }
size_t smem_bytes_strided(uint32_t s, uint32_t log_tile) { return 32ull * ((1u << (s + log_tile)) + (1u << (s ? s - 1 : 0))); }
size_t smem_bytes_last(uint32_t s, uint32_t log_tile) { return 32ull * ((1u << (s + log_tile)) + (1u << (s ? s - 1 : 0))); }

}  // namespace

// Transform of size 2^log_n: reads n_in elements from d_src (zero beyond), leaves the result in d_dst.
// d_src may equal d_work; d_work (2^log_n elements) holds the intermediate passes; d_dst != d_work.
int h2a_ntt_run(h2a_ctx* ctx, const uint8_t* d_src, uint32_t n_in, uint8_t* d_work, uint8_t* d_dst, uint32_t log_n,
                const uint8_t omega[32], int inverse, const uint8_t* coset_shift) {
    if (log_n < 1 || log_n > 28) H2A_FAIL(ctx, H2A_ERR_INVALID, "ntt: log_n=%u not in 1..28", log_n);
    if (!ctx->ntt_attr_set) {
        H2A_CUDA(ctx, cudaFuncSetAttribute(ntt_strided_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        H2A_CUDA(ctx, cudaFuncSetAttribute(ntt_last_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        ctx->ntt_attr_set = true;
    }
    NttTables* tab = nullptr;
    H2A_TRY(get_tables(ctx, log_n, omega, &tab));
    const uint32_t n = 1u << log_n;
    uint32_t s[3];
    int passes;
    split_bits(log_n, s, &passes);

    PassArgs base{};
    base.log_n = log_n;
    base.tw = (const uint8_t*)tab->tw.p;
    base.tw_shift = tab->log_n - log_n;
    base.inverse = inverse;
    base.n_in = n;

    // scaling tables
    namespace hh = h2a_host;
    DevBuf* pre = nullptr;
    DevBuf* post = nullptr;
    int post_mode = 0;
    if (!inverse && coset_shift) {
        H2A_TRY(get_pow_tables(ctx, tab, hh::fr_load(coset_shift), hh::fr_one(), n, &pre));
    } else if (inverse) {
        hh::Fr ninv = hh::inv(hh::fr_from_u64(n));
        if (coset_shift) {
            H2A_TRY(get_pow_tables(ctx, tab, hh::inv(hh::fr_load(coset_shift)), ninv, n, &post));
            post_mode = 2;
        } else {
            H2A_TRY(get_pow_tables(ctx, tab, hh::fr_one(), ninv, 1, &post));
            post_mode = 1;
        }
    }

    h2a_prof_begin(ctx, 1);
    uint32_t done_bits = 0;
    for (int p = 0; p < passes; p++) {
        PassArgs a = base;
        a.s = s[p];
        a.log_r = log_n - done_bits - s[p];
        a.src = (p == 0) ? d_src : d_work;
        a.n_in = (p == 0) ? n_in : n;
        if (p == 0 && pre) {
            a.pre_mode = 1;
            a.pw_lo = (const uint8_t*)pre->p;
            a.pw_hi = a.pw_lo + (32ull << LOG_PW_LO);
        }
        const bool last = (p == passes - 1);
        // measured (tools/sweep.py, H2A_NTT_LOG_TILE): 1024-element tiles (32 KB, 6 blocks per SM by shared memory) beat
        // 2048 from 2^18 up, 512 wins below
        const uint32_t log_tile_elems = log_n <= 17 ? std::min<uint32_t>(9, (uint32_t)ctx->ntt_log_tile) : (uint32_t)ctx->ntt_log_tile;
        uint32_t log_tile_max = log_tile_elems > a.s ? log_tile_elems - a.s : 0;  // tile elements / n_p
        if (!last) {
            a.dst = d_work;
            a.log_tile = std::min(log_tile_max, a.log_r);
            size_t sh = smem_bytes_strided(a.s, a.log_tile);
            ntt_strided_pass_kernel<<<n >> (a.s + a.log_tile), threads_for(a.s + a.log_tile), sh, ctx->stream>>>(a);
        } else {
            a.dst = d_dst;
            a.s1 = passes >= 2 ? s[0] : 0;
            a.s2 = passes >= 3 ? s[1] : 0;
            a.log_tile = std::min(log_tile_max, a.s1);
            if (post) {
                a.post_mode = post_mode;
                a.pw_lo = (const uint8_t*)post->p;
                a.pw_hi = a.pw_lo + (32ull << LOG_PW_LO);
            }
            // TMA staging needs whole rows that are read as they are: no pre-scale, no zero padding, rows of 1 to 16 KB (a 32 KB row is the
            // whole tile: the staging area then halves the blocks per SM and the pass measured 4 % slower, k = 20)
            static const bool bulk_on = !getenv("H2A_NTT_NO_BULK");
            a.bulk = bulk_on && !a.pre_mode && a.n_in == n && a.s >= 5 && a.s <= 9;
            size_t sh = smem_bytes_last(a.s, a.log_tile) + (a.bulk ? 16 + (32ull << (a.s + a.log_tile)) : 0);
            ntt_last_pass_kernel<<<n >> (a.s + a.log_tile), threads_for(a.s + a.log_tile), sh, ctx->stream>>>(a);
        }
        H2A_LAUNCH_CHECK(ctx);
        h2a_prof_mark(ctx);
        done_bits += s[p];
    }
    h2a_prof_end(ctx);
    return H2A_OK;
}

// Public to the other translation units: out <- two-level power tables of `base` scaled by `c`
// (lo at 0: c*base^i for i < 1024; hi at 32*1024: base^(1024*i)).
int h2a_pow_tables(h2a_ctx* ctx, const uint8_t base[32], const uint8_t c[32], uint32_t n, DevBuf& out) {
    return build_pow_tables(ctx, h2a_host::fr_load(base), h2a_host::fr_load(c), n, out);
}
// d_out[i] = c * base^i for i < n  (32-byte elements)
int h2a_pow_vector(h2a_ctx* ctx, const uint8_t base[32], const uint8_t c[32], uint32_t n, uint8_t* d_out) {
    DevBuf pw;
    H2A_TRY(build_pow_tables(ctx, h2a_host::fr_load(base), h2a_host::fr_load(c), n, pw));
    expand_pow_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>((uint8_t*)pw.p, (uint8_t*)pw.p + (32ull << LOG_PW_LO), n, d_out);
    ctx->launches++;
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(pw.p);
    if (e != cudaSuccess) H2A_FAIL(ctx, H2A_ERR_CUDA, "pow_vector: %s", cudaGetErrorString(e));
    return H2A_OK;
}

void h2a_ntt_free_tables(h2a_ctx* ctx) {
    for (auto& kv : ctx->ntt_tables) {
        NttTables* t = kv.second;
        if (t->tw.p) cudaFree(t->tw.p);
        for (auto& pk : t->pow_tables)
            if (pk.second.p) cudaFree(pk.second.p);
        delete t;
    }
    ctx->ntt_tables.clear();
}

const char* h2a_ntt_phase_name(int i) {
    static const char* names[] = {"pass1", "pass2", "pass3"};
    return (i >= 0 && i < 3) ? names[i] : "";
}
