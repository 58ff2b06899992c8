// params.cu — KZG parameter files and the verifier's view of the parameters (row f3 of SURVEY §8).
//
// Replaces, for the drop-in, `Params::write(&mut file)` / `Params::read(file)` (examples/simple-example.rs:679-691: the
// k = 23 parameters are cached under /tmp/halo2-23.params) and `Setup::<Bn256>::verifier_params(&params, public_inputs_size)`
// (:693).  The dependency's own byte format is not visible from the reference (Cargo.toml:12 is an un-vendored branch), so
// the file format is this library's, documented here and in include/h2agg.h; the Rust shim maps Params::{read,write} onto it.
//
//   offset  0   8 bytes  magic "H2APARAM"
//           8   u32      version (1)
//          12   u32      k
//          16   u32      flags: bit 0 = points are compressed to 32 bytes, bit 1 = a 128-byte trailer follows the points
//          20   u32      0
//          24   u64      n = 2^k
//          32   32 bytes 0
//          64   g[0..n), then g_lagrange[0..n):
//                 uncompressed  64 bytes  x || y, each 4 x u64 little-endian Montgomery limbs (the in-memory form)
//                 compressed    32 bytes  canonical x little-endian, bit 255 = parity of canonical y, identity = zeros
//                                         (the proof's point encoding)
//               optional trailer: 128 opaque bytes ([s]G2 for the pairing check; G2 is not on this library's path)
//          end  64 bytes Blake2b-512 (personal "H2A-Params-File\0") of every byte before it
//
// Points stream between the file and HBM in 16 MiB pieces through two pinned buffers (the copy of one piece overlaps the
// read / write of the next); nothing of size n lives in host memory.  On load every point is checked on the device:
// coordinates below p and y^2 = x^3 + 3 (or the identity); a compressed x without a square root is rejected the same way.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "ctx.hpp"
#include "curve.cuh"
#include "host_glue.hpp"

using namespace h2a;

namespace {

constexpr size_t PIECE = 16u << 20;   // bytes per staging buffer

// (p+1)/4: square roots in Fq (p = 3 mod 4)
__device__ __forceinline__ Fq fq_sqrt_candidate(const Fq& a) {
    const uint32_t e[8] = {0xb61f3f52u, 0x4f082305u, 0x5a1c72a3u, 0x65e05aa4u, 0xa0605617u, 0x6e14116du, 0xb84c680au, 0x0c19139cu};
    return a.pow_limbs(e, 252);
}
__device__ __forceinline__ Fq curve_rhs(const Fq& x) {
    const Fq three = Fq::one() + Fq::one() + Fq::one();
    return x.sqr() * x + three;
}

// bad[0] counts points that fail; bad[1] keeps the lowest failing index (g first, then g_lagrange at n + i)
__device__ __forceinline__ void report_bad(unsigned long long* bad, uint64_t i) {
    atomicAdd(bad, 1ull);
    atomicMin(bad + 1, (unsigned long long)i);
}

__global__ void __launch_bounds__(128) params_check_kernel(const uint8_t* __restrict__ pts, uint64_t first, uint32_t n, unsigned long long* bad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Affine p = Affine::load(pts + 64ull * i);
    if (p.is_identity()) return;
    if (Fq::geq_mod(p.x.l) || Fq::geq_mod(p.y.l) || !(p.y.sqr() == curve_rhs(p.x))) report_bad(bad, first + i);
}

// 32-byte compressed -> 64-byte affine (Montgomery)
__global__ void __launch_bounds__(128) params_decompress_kernel(const uint8_t* __restrict__ in, uint64_t first, uint32_t n, uint8_t* __restrict__ out,
                                                                unsigned long long* bad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fq xr = Fq::load(in + 32ull * i);
    const uint32_t sign = xr.l[7] >> 31;
    xr.l[7] &= 0x7fffffffu;
    Affine p;
    if (xr.is_zero() && !sign) {
        p.x = Fq::zero(); p.y = Fq::zero();
    } else if (Fq::geq_mod(xr.l)) {
        report_bad(bad, first + i);
        p.x = Fq::zero(); p.y = Fq::zero();
    } else {
        p.x = xr.to_mont();
        const Fq rhs = curve_rhs(p.x);
        Fq y = fq_sqrt_candidate(rhs);
        if (!(y.sqr() == rhs)) {
            report_bad(bad, first + i);
            p.x = Fq::zero(); y = Fq::zero();
        } else if ((y.from_mont().l[0] & 1u) != sign) {
            y = y.neg();
        }
        p.y = y;
    }
    p.store(out + 64ull * i);
}

// 64-byte affine -> 32-byte compressed
__global__ void __launch_bounds__(128) params_compress_kernel(const uint8_t* __restrict__ in, uint32_t n, uint8_t* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Affine p = Affine::load(in + 64ull * i);
    Fq x = Fq::zero();
    if (!p.is_identity()) {
        x = p.x.from_mont();
        x.l[7] |= (p.y.from_mont().l[0] & 1u) << 31;
    }
    x.store(out + 32ull * i);
}

struct Staging {   // two pinned buffers, two device buffers, released on every exit path
    uint8_t* h[2] = {nullptr, nullptr};
    uint8_t* d[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    unsigned long long* bad = nullptr;
    FILE* f = nullptr;
    ~Staging() {
        for (int i = 0; i < 2; i++) {
            if (h[i]) cudaFreeHost(h[i]);
            if (d[i]) cudaFree(d[i]);
            if (ev[i]) cudaEventDestroy(ev[i]);
        }
        if (bad) cudaFree(bad);
        if (f) fclose(f);
    }
    int init(h2a_ctx* ctx, bool device_bufs) {
        for (int i = 0; i < 2; i++) {
            H2A_CUDA(ctx, cudaHostAlloc((void**)&h[i], PIECE, cudaHostAllocDefault));
            if (device_bufs) H2A_CUDA(ctx, cudaMalloc((void**)&d[i], PIECE));
            H2A_CUDA(ctx, cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
        }
        H2A_CUDA(ctx, cudaMalloc((void**)&bad, 16));
        const unsigned long long init[2] = {0ull, ~0ull};
        H2A_CUDA(ctx, cudaMemcpyAsync(bad, init, 16, cudaMemcpyHostToDevice, ctx->stream));
        return H2A_OK;
    }
};

struct Header {
    char magic[8];
    uint32_t version, k, flags, zero;
    uint64_t n;
    uint8_t reserved[32];
};
static_assert(sizeof(Header) == 64, "params header is 64 bytes");
const char MAGIC[8] = {'H', '2', 'A', 'P', 'A', 'R', 'A', 'M'};
const char PERSONAL[16] = {'H', '2', 'A', '-', 'P', 'a', 'r', 'a', 'm', 's', '-', 'F', 'i', 'l', 'e', 0};

}  // namespace

extern "C" {

static int params_write_impl(h2a_ctx* ctx, const char* path, uint32_t k, const h2a_bases* g, const h2a_bases* g_lagrange, int compressed,
                             const uint8_t* trailer128);
int h2a_params_write(h2a_ctx* ctx, const char* path, uint32_t k, const h2a_bases* g, const h2a_bases* g_lagrange, int compressed,
                     const uint8_t* trailer128) {
    H2A_DEVICE(ctx);
    const int rc = params_write_impl(ctx, path, k, g, g_lagrange, compressed, trailer128);
    if (rc != H2A_OK && path && ctx && ctx->err.find("cannot create") == std::string::npos) remove(path);   // no truncated file is left behind
    return rc;
}
static int params_write_impl(h2a_ctx* ctx, const char* path, uint32_t k, const h2a_bases* g, const h2a_bases* g_lagrange, int compressed,
                             const uint8_t* trailer128) {
    if (!ctx || !path || !g || !g_lagrange) return H2A_ERR_INVALID;
    if (k < 1 || k > 26) H2A_FAIL(ctx, H2A_ERR_INVALID, "params_write: k=%u not in 1..26", k);
    const uint64_t n = 1ull << k;
    if (g->n != n || g_lagrange->n != n) H2A_FAIL(ctx, H2A_ERR_INVALID, "params_write: handles hold %zu / %zu points, 2^k = %llu", g->n, g_lagrange->n, (unsigned long long)n);
    Staging s;
    H2A_TRY(s.init(ctx, compressed != 0));
    s.f = fopen(path, "wb");
    if (!s.f) H2A_FAIL(ctx, H2A_ERR_INVALID, "params_write: cannot create %s", path);
    h2a_glue::Blake2bState hash(64, PERSONAL);
    Header hd;
    memset(&hd, 0, sizeof hd);
    memcpy(hd.magic, MAGIC, 8);
    hd.version = 1; hd.k = k; hd.flags = (compressed ? 1u : 0u) | (trailer128 ? 2u : 0u); hd.n = n;
    hash.absorb((const uint8_t*)&hd, sizeof hd);
    if (fwrite(&hd, sizeof hd, 1, s.f) != 1) H2A_FAIL(ctx, H2A_ERR_INVALID, "params_write: write to %s failed", path);
    const size_t out_elem = compressed ? 32 : 64, per = PIECE / 64;   // points per piece
    for (const h2a_bases* b : {g, g_lagrange}) {
        // piece i is copied (and compressed) into buffer i & 1 while piece i - 1 is hashed and written
        size_t queued = 0, written = 0;
        size_t cnt[2] = {0, 0};
        while (written < n) {
            while (queued < n && queued - written < 2 * per) {   // the buffer of piece i is free once piece i - 2 is written
                const int sl = (int)((queued / per) & 1);
                cnt[sl] = (size_t)std::min<uint64_t>(per, n - queued);
                if (compressed) {
                    params_compress_kernel<<<(unsigned)((cnt[sl] + 127) / 128), 128, 0, ctx->stream>>>(b->d + 64 * queued, (uint32_t)cnt[sl], s.d[sl]);
                    H2A_LAUNCH_CHECK(ctx);
                    H2A_CUDA(ctx, cudaMemcpyAsync(s.h[sl], s.d[sl], cnt[sl] * 32, cudaMemcpyDeviceToHost, ctx->stream));
                } else {
                    H2A_CUDA(ctx, cudaMemcpyAsync(s.h[sl], b->d + 64 * queued, cnt[sl] * 64, cudaMemcpyDeviceToHost, ctx->stream));
                }
                H2A_CUDA(ctx, cudaEventRecord(s.ev[sl], ctx->stream));
                queued += cnt[sl];
            }
            const int sl = (int)((written / per) & 1);
            H2A_CUDA(ctx, cudaEventSynchronize(s.ev[sl]));
            hash.absorb(s.h[sl], cnt[sl] * out_elem);
            if (fwrite(s.h[sl], out_elem, cnt[sl], s.f) != cnt[sl]) H2A_FAIL(ctx, H2A_ERR_INVALID, "params_write: write to %s failed", path);
            written += cnt[sl];
        }
    }
    if (trailer128) {
        hash.absorb(trailer128, 128);
        if (fwrite(trailer128, 128, 1, s.f) != 1) H2A_FAIL(ctx, H2A_ERR_INVALID, "params_write: write to %s failed", path);
    }
    uint8_t digest[64];
    hash.digest(digest);
    if (fwrite(digest, 64, 1, s.f) != 1 || fflush(s.f) != 0) H2A_FAIL(ctx, H2A_ERR_INVALID, "params_write: write to %s failed", path);
    return H2A_OK;
}

int h2a_params_read(h2a_ctx* ctx, const char* path, uint32_t* out_k, h2a_bases** out_g, h2a_bases** out_g_lagrange, uint8_t* trailer128,
                    int* has_trailer) {
    H2A_DEVICE(ctx);
    if (!ctx || !path || !out_k || !out_g || !out_g_lagrange) return H2A_ERR_INVALID;
    *out_g = *out_g_lagrange = nullptr;
    if (has_trailer) *has_trailer = 0;
    Staging s;
    H2A_TRY(s.init(ctx, true));
    s.f = fopen(path, "rb");
    if (!s.f) H2A_FAIL(ctx, H2A_ERR_INVALID, "params_read: cannot open %s", path);
    h2a_glue::Blake2bState hash(64, PERSONAL);
    Header hd;
    if (fread(&hd, sizeof hd, 1, s.f) != 1) H2A_FAIL(ctx, H2A_ERR_INVALID, "params_read: %s is shorter than its header", path);
    if (memcmp(hd.magic, MAGIC, 8) != 0) H2A_FAIL(ctx, H2A_ERR_INVALID, "params_read: %s is not a parameter file (magic)", path);
    if (hd.version != 1 || hd.k < 1 || hd.k > 26 || hd.n != (1ull << hd.k) || (hd.flags & ~3u))
        H2A_FAIL(ctx, H2A_ERR_INVALID, "params_read: unsupported header (version %u, k %u, flags %u)", hd.version, hd.k, hd.flags);
    hash.absorb((const uint8_t*)&hd, sizeof hd);
    const bool compressed = hd.flags & 1u;
    const uint64_t n = hd.n;
    const size_t in_elem = compressed ? 32 : 64, per = PIECE / 64;
    uint8_t* dev[2] = {nullptr, nullptr};
    auto drop = [&]() { for (uint8_t* q : dev) if (q) cudaFree(q); };
    for (int which = 0; which < 2; which++) {
        cudaError_t e = cudaMalloc((void**)&dev[which], 64 * n);
        if (e != cudaSuccess) { drop(); H2A_FAIL(ctx, H2A_ERR_OOM, "params_read: cudaMalloc(%llu): %s", (unsigned long long)(64 * n), cudaGetErrorString(e)); }
    }
#define PR_FAIL(code, ...) do { cudaStreamSynchronize(ctx->stream); drop(); H2A_FAIL(ctx, code, __VA_ARGS__); } while (0)
#define PR_CUDA(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) PR_FAIL(H2A_ERR_CUDA, "params_read: %s -> %s", #call, cudaGetErrorString(_e)); } while (0)
    for (int which = 0; which < 2; which++) {
        size_t done = 0, piece = 0;
        while (done < n) {
            const int sl = (int)(piece & 1);
            const size_t cnt = (size_t)std::min<uint64_t>(per, n - done);
            if (piece >= 2) PR_CUDA(cudaEventSynchronize(s.ev[sl]));    // the copy that last used this pinned buffer is over
            if (fread(s.h[sl], in_elem, cnt, s.f) != cnt) PR_FAIL(H2A_ERR_INVALID, "params_read: %s is truncated", path);
            hash.absorb(s.h[sl], cnt * in_elem);
            uint8_t* dst = dev[which] + 64 * done;
            if (compressed) {
                PR_CUDA(cudaMemcpyAsync(s.d[sl], s.h[sl], cnt * 32, cudaMemcpyHostToDevice, ctx->stream));
                params_decompress_kernel<<<(unsigned)((cnt + 127) / 128), 128, 0, ctx->stream>>>(s.d[sl], (uint64_t)which * n + done, (uint32_t)cnt, dst, s.bad);
            } else {
                PR_CUDA(cudaMemcpyAsync(dst, s.h[sl], cnt * 64, cudaMemcpyHostToDevice, ctx->stream));
                params_check_kernel<<<(unsigned)((cnt + 127) / 128), 128, 0, ctx->stream>>>(dst, (uint64_t)which * n + done, (uint32_t)cnt, s.bad);
            }
            ctx->launches++;
            PR_CUDA(cudaGetLastError());
            PR_CUDA(cudaEventRecord(s.ev[sl], ctx->stream));
            done += cnt;
            piece++;
        }
        PR_CUDA(cudaStreamSynchronize(ctx->stream));   // both pinned buffers are free again for the next array
    }
    if (hd.flags & 2u) {
        uint8_t tr[128];
        if (fread(tr, 128, 1, s.f) != 1) PR_FAIL(H2A_ERR_INVALID, "params_read: %s is truncated (trailer)", path);
        hash.absorb(tr, 128);
        if (trailer128) memcpy(trailer128, tr, 128);
        if (has_trailer) *has_trailer = 1;
    }
    uint8_t want[64], got[64], extra;
    hash.digest(want);
    if (fread(got, 64, 1, s.f) != 1) PR_FAIL(H2A_ERR_INVALID, "params_read: %s is truncated (digest)", path);
    if (fread(&extra, 1, 1, s.f) == 1) PR_FAIL(H2A_ERR_INVALID, "params_read: %s has trailing bytes", path);
    if (memcmp(want, got, 64) != 0) PR_FAIL(H2A_ERR_INVALID, "params_read: %s fails its Blake2b digest (corrupt file)", path);
    unsigned long long bad[2];
    PR_CUDA(cudaMemcpy(bad, s.bad, 16, cudaMemcpyDeviceToHost));
    if (bad[0]) PR_FAIL(H2A_ERR_INVALID, "params_read: %llu point(s) of %s are not on the curve (first: %s[%llu])", bad[0], path,
                        bad[1] < n ? "g" : "g_lagrange", bad[1] % n);
#undef PR_CUDA
#undef PR_FAIL
    h2a_bases *bg = new h2a_bases(), *bl = new h2a_bases();
    bg->d = dev[0]; bg->n = n; bg->owned = true;
    bl->d = dev[1]; bl->n = n; bl->owned = true;
    *out_k = hd.k;
    *out_g = bg;
    *out_g_lagrange = bl;
    return H2A_OK;
}

// `Setup::verifier_params(&params, public_inputs_size)`: the verifier commits its public inputs against the first
// `public_inputs_size` Lagrange bases only (examples/simple-example.rs:590,693 and :638-640).  The view shares the
// device memory of `g_lagrange` (which must outlive it) and owns nothing; free it with h2a_bases_free.
int h2a_params_verifier_view(h2a_ctx* ctx, const h2a_bases* g_lagrange, size_t public_inputs_size, h2a_bases** out) {
    H2A_DEVICE(ctx);
    if (!ctx || !g_lagrange || !out) return H2A_ERR_INVALID;
    if (public_inputs_size > g_lagrange->n) H2A_FAIL(ctx, H2A_ERR_INVALID, "verifier_view: %zu public inputs exceed the %zu Lagrange bases", public_inputs_size, g_lagrange->n);
    h2a_bases* b = new h2a_bases();
    b->d = g_lagrange->d;
    b->n = public_inputs_size;
    b->owned = false;
    *out = b;
    return H2A_OK;
}

}  // extern "C"
