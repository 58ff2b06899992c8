// misc.cu — context lifecycle, workspace, profiling marks, synthetic-input generators,
// element-wise test hooks and the integer-pipe micro-benchmarks of include/h2agg.h.
#include <cstdlib>
#include <cstring>

#include "ctx.hpp"
#include "curve.cuh"

using namespace h2a;

const char* h2a_msm_phase_name(int i);
const char* h2a_ntt_phase_name(int i);
void h2a_ntt_free_tables(h2a_ctx* ctx);

// ------------------------------------------------------------------ workspace / profiling
int h2a_reserve(h2a_ctx* ctx, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return H2A_OK;
    if (b.p) {
        H2A_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        H2A_CUDA(ctx, cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = bytes + bytes / 8 + 256;
    H2A_CUDA(ctx, cudaMalloc(&b.p, want));
    b.cap = want;
    return H2A_OK;
}
int h2a_reserve_pinned(h2a_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->pinned_cap) return H2A_OK;
    if (ctx->pinned) {
        H2A_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        H2A_CUDA(ctx, cudaFreeHost(ctx->pinned));
        ctx->pinned = nullptr;
        ctx->pinned_cap = 0;
    }
    size_t want = std::max<size_t>(bytes, 1 << 16);
    H2A_CUDA(ctx, cudaMallocHost(&ctx->pinned, want));
    ctx->pinned_cap = want;
    return H2A_OK;
}
void h2a_prof_begin(h2a_ctx* ctx, int kind) {
    if (!ctx->profiling) return;
    ctx->last_kind = kind;
    ctx->ev_used = 0;
    h2a_prof_mark(ctx);
}
void h2a_prof_mark(h2a_ctx* ctx) {
    if (!ctx->profiling) return;
    if (ctx->ev_used == (int)ctx->ev.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        ctx->ev.push_back(e);
    }
    cudaEventRecord(ctx->ev[ctx->ev_used++], ctx->stream);
}
void h2a_prof_end(h2a_ctx* ctx) {
    if (!ctx->profiling) return;
    cudaStreamSynchronize(ctx->stream);
    ctx->phase_ms.clear();
    for (int i = 1; i < ctx->ev_used; i++) {
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev[i - 1], ctx->ev[i]);
        ctx->phase_ms.push_back(ms);
    }
}

// ------------------------------------------------------------------ kernels
namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
// 254 uniform bits for (seed, stream, i, attempt) as 8 x u32 limbs; returns spare bits
__device__ __forceinline__ uint64_t draw254(uint64_t seed, uint64_t stream, uint64_t i, uint64_t attempt,
                                            uint32_t out[8]) {
    uint64_t h = mix64(seed + 0x100000001b3ull * stream);
    h = mix64(h ^ i);
    h = mix64(h + attempt);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        uint64_t v = mix64(h + (uint64_t)j + 1);
        if (j == 3) v &= 0x3fffffffffffffffull;
        out[2 * j] = (uint32_t)v;
        out[2 * j + 1] = (uint32_t)(v >> 32);
    }
    return mix64(h + 5);
}

__global__ void gen_scalars_kernel(uint64_t seed, uint64_t first, uint64_t n, uint8_t* out) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr v;
    for (uint64_t att = 0;; att++) {
        draw254(seed, 1, first + i, att, v.l);
        if (!Fr::geq_mod(v.l)) break;
    }
    v.store(out + 32 * i);  // the drawn value is the in-memory (Montgomery) form
}

__global__ void __launch_bounds__(128) gen_bases_kernel(uint64_t seed, uint64_t first, uint64_t n, uint8_t* out) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    // (p+1)/4
    const uint32_t e[8] = {0xb61f3f52u, 0x4f082305u, 0x5a1c72a3u, 0x65e05aa4u,
                           0xa0605617u, 0x6e14116du, 0xb84c680au, 0x0c19139cu};
    Fq three = Fq::one() + Fq::one() + Fq::one();
    for (uint64_t att = 0;; att++) {
        Fq xr;
        uint64_t spare = draw254(seed, 2, first + i, att, xr.l);
        if (Fq::geq_mod(xr.l)) continue;
        Fq x = xr.to_mont();
        Fq rhs = x.sqr() * x + three;
        Fq y = rhs.pow_limbs(e, 252);
        if (!(y.sqr() == rhs)) continue;
        if (spare & 1) y = y.neg();
        x.store(out + 64 * i);
        y.store(out + 64 * i + 32);
        return;
    }
}

template <int F>
__global__ void field_op_kernel(int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    typedef Fp<F> T;
    T x = T::load(a + 32 * i), y = b ? T::load(b + 32 * i) : T::zero(), r;
    switch (op) {
        case 0: r = x + y; break;
        case 1: r = x - y; break;
        case 2: r = x * y; break;
        case 3: r = x.sqr(); break;
        case 4: r = x.inv_fermat(); break;
        case 6: r = x.to_mont(); break;
        case 7: r = x.from_mont(); break;
        case 8: r = x.inv(); break;
        default: r = x.neg(); break;
    }
    r.store(out + 32 * i);
}

__device__ Affine xyzz_to_affine(const XYZZ& p) {
    Affine r;
    if (p.is_identity()) {
        r.x = Fq::zero();
        r.y = Fq::zero();
        return r;
    }
    Fq zi = p.zzz.inv();
    Fq zzi = (zi * p.zz).sqr();
    r.x = p.x * zzi;
    r.y = p.y * zi;
    return r;
}

__global__ void __launch_bounds__(128) g1_op_kernel(int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine pa = Affine::load(a + 64 * i);
    XYZZ acc = XYZZ::from_affine(pa);
    if (op == 0) {
        Affine pb = Affine::load(b + 64 * i);
        if (!pb.is_identity()) acc.add_affine(pb, false);
    } else if (op == 1) {
        acc = acc.dbl();
    } else {  // op 2: full XYZZ add of (2a) and b, exercising XYZZ::add with non-unit ZZ on one side
        Affine pb = Affine::load(b + 64 * i);
        XYZZ d = acc.dbl();
        XYZZ q = XYZZ::from_affine(pb);
        d.add(q);
        acc = d;
    }
    xyzz_to_affine(acc).store(out + 64 * i);
}

// dependent-free 32-bit IMAD chains: 8 independent accumulators per thread
__global__ void __launch_bounds__(256) imad_bench_kernel(uint32_t* out, int iters, uint32_t a, uint32_t b) {
    uint32_t x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int k = 0; k < 8; k++) x[k] = x[k] * a + b;
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) modmul_bench_kernel(uint8_t* out, int iters) {
    Fq a = Fq::one(), b = Fq::r2();
    a.l[0] ^= threadIdx.x;
    Fq c = a + b, d = b + b;
    for (int i = 0; i < iters; i++) {
        a = a * b;
        c = c * d;
        b = b * a;
        d = d * c;
    }
    (a + b + c + d).store(out + 32ull * (blockIdx.x * blockDim.x + threadIdx.x));
}

}  // namespace

// ------------------------------------------------------------------ C ABI
extern "C" {

int h2a_version(void) { return 1; }

int h2a_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int h2a_init(h2a_ctx** out, int device) {
    if (!out) return H2A_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device < 0 || device >= n) return H2A_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return H2A_ERR_NO_DEVICE;
    if (prop.major != 10) return H2A_ERR_NO_DEVICE;  // sm_100a cubin only
    if (cudaSetDevice(device) != cudaSuccess) return H2A_ERR_CUDA;
    h2a_ctx* ctx = new h2a_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    // the lanes that carry commitments run at the highest stream priority: work queued on a default-priority stream
    // (the prover's transform lane, plonk_prove.cu) then only fills the SMs they leave idle
    int prio_least = 0, prio_greatest = 0;
    cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
    ctx->stream_priority = prio_greatest;
    if (cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, ctx->stream_priority) != cudaSuccess) {
        delete ctx;
        return H2A_ERR_CUDA;
    }
    const char* env = getenv("H2A_MSM_WINDOW");
    if (env) ctx->msm_window_override = atoi(env);
    env = getenv("H2A_NTT_LOG_TILE");
    if (env && atoi(env) >= 8 && atoi(env) <= 12) ctx->ntt_log_tile = atoi(env);
    env = getenv("H2A_MSM_SEG");
    if (env && atoi(env) >= 1 && atoi(env) <= 64) ctx->msm_seg_len = atoi(env);
    env = getenv("H2A_MSM_RED_CHUNK");
    if (env && atoi(env) >= 128 && atoi(env) <= 65536 && (atoi(env) & (atoi(env) - 1)) == 0) ctx->msm_red_chunk = atoi(env);
    env = getenv("H2A_MSM_HOST_SPLIT");
    if (env && atoi(env) >= 1 && atoi(env) <= 16) ctx->msm_host_split = atoi(env);
    env = getenv("H2A_MSM_GROUP");
    if (env && atoi(env) >= 1 && atoi(env) <= 64) ctx->msm_group_cols = atoi(env);
    env = getenv("H2A_MSM_GROUP_HOST");
    if (env && atoi(env) >= 1 && atoi(env) <= 64) ctx->msm_group_cols_host = atoi(env);
    env = getenv("H2A_MSM_ROUNDS_BIAS");
    if (env && atoi(env) >= -3 && atoi(env) <= 3) ctx->msm_rounds_bias = atoi(env);
    env = getenv("H2A_MSM_TREE_PRIO");
    if (env) ctx->msm_tree_prio = atoi(env) ? 1 : 0;
    env = getenv("H2A_MSM_ALGO");
    if (env) ctx->msm_algo = atoi(env) ? 1 : 0;
    *out = ctx;
    return H2A_OK;
}

int h2a_destroy(h2a_ctx* ctx) {
    H2A_DEVICE(ctx);
    if (!ctx) return H2A_ERR_INVALID;
    if (ctx->alt) {
        h2a_destroy(ctx->alt);
        ctx->alt = nullptr;
    }
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    h2a_comm_destroy(ctx);
    DevBuf* bufs[] = {&ctx->comm_buf, &ctx->scalars, &ctx->offsets, &ctx->cursor, &ctx->sorted, &ctx->buckets, &ctx->segsums,
                      &ctx->winsums, &ctx->heavy,   &ctx->misc,   &ctx->ntt_a,  &ctx->ntt_b,  &ctx->aff_a, &ctx->aff_b, &ctx->aff_c,
                      &ctx->aff_scratch, &ctx->aff_u32};
    for (DevBuf* b : bufs)
        if (b->p) cudaFree(b->p);
    h2a_ntt_free_tables(ctx);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    for (cudaEvent_t e : ctx->ev) cudaEventDestroy(e);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    for (int h = 0; h < 2; h++) {
        if (ctx->stream_lo[h]) cudaStreamDestroy(ctx->stream_lo[h]);
        for (int e = 0; e < 2; e++) if (ctx->ev_tree[h][e]) cudaEventDestroy(ctx->ev_tree[h][e]);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return H2A_OK;
}

}  // extern "C"

int h2a_get_alt(h2a_ctx* ctx, h2a_ctx** out) {
    if (!ctx->alt) {
        int rc = h2a_init(&ctx->alt, ctx->device);
        if (rc != H2A_OK) H2A_FAIL(ctx, rc, "could not create the second lane");
    }
    *out = ctx->alt;
    return H2A_OK;
}

extern "C" {

const char* h2a_last_error(const h2a_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }
void* h2a_stream(h2a_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int h2a_sync(h2a_ctx* ctx) {
    H2A_DEVICE(ctx);
    if (!ctx) return H2A_ERR_INVALID;
    H2A_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return H2A_OK;
}
int h2a_dev_alloc(h2a_ctx* ctx, size_t bytes, void** out_dev) {
    H2A_DEVICE(ctx);
    if (!ctx || !out_dev) return H2A_ERR_INVALID;
    H2A_CUDA(ctx, cudaSetDevice(ctx->device));
    H2A_CUDA(ctx, cudaMalloc(out_dev, bytes ? bytes : 1));
    return H2A_OK;
}
int h2a_dev_free(h2a_ctx* ctx, void* dev) {
    H2A_DEVICE(ctx);
    if (!ctx) return H2A_ERR_INVALID;
    H2A_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    H2A_CUDA(ctx, cudaFree(dev));
    return H2A_OK;
}
int h2a_copy_h2d(h2a_ctx* ctx, void* dev, const void* host, size_t bytes) {
    H2A_DEVICE(ctx);
    if (!ctx || ((!dev || !host) && bytes)) return H2A_ERR_INVALID;
    if (!bytes) return H2A_OK;
    H2A_CUDA(ctx, cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    H2A_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return H2A_OK;
}
int h2a_copy_d2h(h2a_ctx* ctx, void* host, const void* dev, size_t bytes) {
    H2A_DEVICE(ctx);
    if (!ctx || ((!dev || !host) && bytes)) return H2A_ERR_INVALID;
    if (!bytes) return H2A_OK;
    H2A_CUDA(ctx, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    H2A_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return H2A_OK;
}

int h2a_gen_scalars_dev(h2a_ctx* ctx, uint64_t seed, size_t first, size_t n, void* d_out) {
    H2A_DEVICE(ctx);
    if (!ctx || (!d_out && n)) return H2A_ERR_INVALID;
    if (!n) return H2A_OK;
    gen_scalars_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(seed, first, n, (uint8_t*)d_out);
    H2A_LAUNCH_CHECK(ctx);
    H2A_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return H2A_OK;
}
int h2a_gen_bases_dev(h2a_ctx* ctx, uint64_t seed, size_t first, size_t n, void* d_out) {
    H2A_DEVICE(ctx);
    if (!ctx || (!d_out && n)) return H2A_ERR_INVALID;
    if (!n) return H2A_OK;
    gen_bases_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(seed, first, n, (uint8_t*)d_out);
    H2A_LAUNCH_CHECK(ctx);
    H2A_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return H2A_OK;
}

namespace {
struct ScopedDev {  // device allocation released on every exit path
    void* p = nullptr;
    ~ScopedDev() { if (p) cudaFree(p); }
};
}  // namespace

static int run_elementwise(h2a_ctx* ctx, int kind, int field, int op, const uint8_t* a, const uint8_t* b, uint8_t* out,
                           size_t n, size_t elem) {
    if (!ctx || !a || !out) return H2A_ERR_INVALID;
    if (!n) return H2A_OK;
    ScopedDev da, db, dout;
    H2A_CUDA(ctx, cudaMalloc(&da.p, n * elem));
    H2A_CUDA(ctx, cudaMalloc(&dout.p, n * elem));
    if (b) H2A_CUDA(ctx, cudaMalloc(&db.p, n * elem));
    H2A_CUDA(ctx, cudaMemcpyAsync(da.p, a, n * elem, cudaMemcpyHostToDevice, ctx->stream));
    if (b) H2A_CUDA(ctx, cudaMemcpyAsync(db.p, b, n * elem, cudaMemcpyHostToDevice, ctx->stream));
    unsigned blocks = (unsigned)((n + 127) / 128);
    if (kind == 0) {
        if (field == 0)
            field_op_kernel<FQ><<<blocks, 128, 0, ctx->stream>>>(op, (uint8_t*)da.p, (uint8_t*)db.p, (uint8_t*)dout.p, n);
        else
            field_op_kernel<FR><<<blocks, 128, 0, ctx->stream>>>(op, (uint8_t*)da.p, (uint8_t*)db.p, (uint8_t*)dout.p, n);
    } else {
        g1_op_kernel<<<blocks, 128, 0, ctx->stream>>>(op, (uint8_t*)da.p, (uint8_t*)db.p, (uint8_t*)dout.p, n);
    }
    H2A_LAUNCH_CHECK(ctx);
    H2A_CUDA(ctx, cudaMemcpyAsync(out, dout.p, n * elem, cudaMemcpyDeviceToHost, ctx->stream));
    H2A_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return H2A_OK;
}
int h2a_field_op(h2a_ctx* ctx, int field, int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
    H2A_DEVICE(ctx);
    if (field < 0 || field > 1 || op < 0 || op > 8) return H2A_ERR_INVALID;
    if (op <= 2 && !b) return H2A_ERR_INVALID;
    return run_elementwise(ctx, 0, field, op, a, op <= 2 ? b : nullptr, out, n, 32);
}
int h2a_g1_op(h2a_ctx* ctx, int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
    H2A_DEVICE(ctx);
    if (op < 0 || op > 2) return H2A_ERR_INVALID;
    if (op != 1 && !b) return H2A_ERR_INVALID;
    return run_elementwise(ctx, 1, 0, op, a, op != 1 ? b : nullptr, out, n, 64);
}

int h2a_set_profiling(h2a_ctx* ctx, int on) {
    if (!ctx) return H2A_ERR_INVALID;
    ctx->profiling = on != 0;
    return H2A_OK;
}
int h2a_last_phase_ms(h2a_ctx* ctx, float* ms, int cap) {
    if (!ctx || !ms) return H2A_ERR_INVALID;
    int n = std::min<int>(cap, (int)ctx->phase_ms.size());
    for (int i = 0; i < n; i++) ms[i] = ctx->phase_ms[i];
    return n;
}
const char* h2a_phase_name(int kind, int index) { return kind == 0 ? h2a_msm_phase_name(index) : h2a_ntt_phase_name(index); }
uint64_t h2a_launch_count(const h2a_ctx* ctx) { return ctx ? ctx->launches : 0; }

int h2a_bench_imad(h2a_ctx* ctx, double* out) {
    H2A_DEVICE(ctx);
    if (!ctx || !out) return H2A_ERR_INVALID;
    const int blocks = ctx->sm_count * 8, threads = 256, iters = 4096;
    H2A_TRY(h2a_reserve(ctx, ctx->misc, (size_t)blocks * threads * 4));
    cudaEvent_t e0, e1;
    H2A_CUDA(ctx, cudaEventCreate(&e0));
    H2A_CUDA(ctx, cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        H2A_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
        imad_bench_kernel<<<blocks, threads, 0, ctx->stream>>>((uint32_t*)ctx->misc.p, iters, 0x9e3779b1u + rep, 12345u);
        H2A_LAUNCH_CHECK(ctx);
        H2A_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        H2A_CUDA(ctx, cudaEventSynchronize(e1));
        float ms;
        H2A_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best = std::min(best, ms);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    double ops = (double)blocks * threads * iters * 64.0;
    *out = ops / (best * 1e-3) / 1e12;
    return H2A_OK;
}
int h2a_bench_modmul(h2a_ctx* ctx, double* out) {
    H2A_DEVICE(ctx);
    if (!ctx || !out) return H2A_ERR_INVALID;
    const int blocks = ctx->sm_count * 8, threads = 256, iters = 2048;
    H2A_TRY(h2a_reserve(ctx, ctx->misc, (size_t)blocks * threads * 32));
    cudaEvent_t e0, e1;
    H2A_CUDA(ctx, cudaEventCreate(&e0));
    H2A_CUDA(ctx, cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        H2A_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
        modmul_bench_kernel<<<blocks, threads, 0, ctx->stream>>>((uint8_t*)ctx->misc.p, iters);
        H2A_LAUNCH_CHECK(ctx);
        H2A_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        H2A_CUDA(ctx, cudaEventSynchronize(e1));
        float ms;
        H2A_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best = std::min(best, ms);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    double ops = (double)blocks * threads * iters * 4.0;
    *out = ops / (best * 1e-3) / 1e9;
    return H2A_OK;
}

}  // extern "C"
