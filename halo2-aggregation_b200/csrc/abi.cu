// abi.cu — the MSM / NTT entry points of include/h2agg.h: argument checks, host<->device staging,
// and dispatch to the kernels in msm.cu / ntt.cu.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ctx.hpp"
#include "host_bn254.hpp"
#include "dist_layout.hpp"
#include "tree_layout.hpp"

int h2a_ntt_run(h2a_ctx* ctx, const uint8_t* d_src, uint32_t n_in, uint8_t* d_work, uint8_t* d_dst, uint32_t log_n,
                const uint8_t omega[32], int inverse, const uint8_t* coset_shift);
int h2a_small_msm(h2a_ctx* ctx, const uint8_t* bases, const uint8_t* scalars, const uint32_t* sum_offsets, size_t n_sums,
                  uint8_t* out_affine);

constexpr size_t SMALL_MSM_MAX = 2048;  // below this an MSM is one block of the small-MSM kernel

extern "C" {

// ------------------------------------------------------------------ bases
int h2a_bases_upload(h2a_ctx* ctx, const uint8_t* affine_xy, size_t n, h2a_bases** out) {
    H2A_DEVICE(ctx);
    if (!ctx || !out || (!affine_xy && n)) return H2A_ERR_INVALID;
    h2a_bases* b = new h2a_bases();
    void* d = nullptr;
    cudaError_t e = cudaMalloc(&d, n ? n * 64 : 64);
    if (e != cudaSuccess) {
        delete b;
        H2A_FAIL(ctx, H2A_ERR_OOM, "bases_upload: cudaMalloc(%zu): %s", n * 64, cudaGetErrorString(e));
    }
    if (n) {
        e = cudaMemcpyAsync(d, affine_xy, n * 64, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            cudaFree(d);
            delete b;
            H2A_FAIL(ctx, H2A_ERR_CUDA, "bases_upload: copy: %s", cudaGetErrorString(e));
        }
    }
    b->d = (const uint8_t*)d;
    b->n = n;
    b->owned = true;
    *out = b;
    return H2A_OK;
}
int h2a_bases_from_device(h2a_ctx* ctx, const void* d_affine_xy, size_t n, h2a_bases** out) {
    H2A_DEVICE(ctx);
    if (!ctx || !out || (!d_affine_xy && n)) return H2A_ERR_INVALID;
    h2a_bases* b = new h2a_bases();
    b->d = (const uint8_t*)d_affine_xy;
    b->n = n;
    b->owned = false;
    *out = b;
    return H2A_OK;
}
int h2a_bases_free(h2a_ctx* ctx, h2a_bases* bases) {
    H2A_DEVICE(ctx);
    if (!ctx || !bases) return H2A_ERR_INVALID;
    H2A_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (bases->table) H2A_CUDA(ctx, cudaFree(bases->table));
    if (bases->owned && bases->d) H2A_CUDA(ctx, cudaFree((void*)bases->d));
    delete bases;
    return H2A_OK;
}
size_t h2a_bases_len(const h2a_bases* bases) { return bases ? bases->n : 0; }
int h2a_bases_download(h2a_ctx* ctx, const h2a_bases* bases, uint8_t* out_affine_xy) {
    H2A_DEVICE(ctx);
    if (!ctx || !bases || (!out_affine_xy && bases->n)) return H2A_ERR_INVALID;
    if (!bases->n) return H2A_OK;
    H2A_CUDA(ctx, cudaMemcpyAsync(out_affine_xy, bases->d, 64 * bases->n, cudaMemcpyDeviceToHost, ctx->stream));
    H2A_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return H2A_OK;
}
int h2a_bases_precompute(h2a_ctx* ctx, h2a_bases* bases, int window_bits) {
    H2A_DEVICE(ctx);
    if (!ctx || !bases) return H2A_ERR_INVALID;
    if (window_bits == 0) {  // drop the tables
        H2A_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (bases->table) H2A_CUDA(ctx, cudaFree(bases->table));
        bases->table = nullptr;
        bases->table_c = 0;
        return H2A_OK;
    }
    return h2a_msm_precompute(ctx, bases, window_bits);
}

// ------------------------------------------------------------------ MSM
int h2a_msm_set_window(h2a_ctx* ctx, int c) {
    if (!ctx || (c != 0 && (c < 6 || c > 20))) return H2A_ERR_INVALID;
    ctx->msm_window_override = c;
    return H2A_OK;
}

int h2a_msm_set_algorithm(h2a_ctx* ctx, int algo) {
    if (!ctx || algo < 0 || algo > 1) return H2A_ERR_INVALID;
    ctx->msm_algo = algo;
    return H2A_OK;
}

int h2a_msm_set_host_split(h2a_ctx* ctx, int pieces) {
    if (!ctx || pieces < 1 || pieces > 16) return H2A_ERR_INVALID;
    ctx->msm_host_split = pieces;
    return H2A_OK;
}

int h2a_msm_set_group(h2a_ctx* ctx, int cols, int cols_host) {
    if (!ctx || cols < 1 || cols > 64 || cols_host < 1 || cols_host > 64) return H2A_ERR_INVALID;
    ctx->msm_group_cols = cols;
    ctx->msm_group_cols_host = cols_host;
    return H2A_OK;
}

int h2a_tree_layout(uint64_t total_padded, int rounds, int half, int round, int64_t out6[6]) {
    if (!out6 || rounds < 1 || rounds > 5 || half < 0 || half > 1 || round < 0 || round >= rounds) return H2A_ERR_INVALID;
    if (total_padded == 0 || total_padded % ((uint64_t)2 << rounds)) return H2A_ERR_INVALID;   // msm.cu rounds it up to 2^(R+1)
    TreeSpan in, out;
    tree_round_spans(total_padded, rounds, half, round, &in, &out);
    out6[0] = in.array; out6[1] = (int64_t)in.first; out6[2] = (int64_t)in.count;
    out6[3] = out.array; out6[4] = (int64_t)out.first; out6[5] = (int64_t)out.count;
    return H2A_OK;
}

int h2a_dist_window(uint32_t m, int world, int rank, uint32_t halo, uint32_t out4[4]) {
    if (!out4 || world < 2 || rank < 0 || rank >= world || m == 0 || m % (uint32_t)world || 2ull * halo > m / (uint32_t)world) return H2A_ERR_INVALID;
    DistPiece piece[2] = {{0, 0}, {0, 0}};
    const int n = dist_window(m, world, rank, halo, piece);
    out4[0] = piece[0].first; out4[1] = piece[0].count; out4[2] = piece[1].first; out4[3] = piece[1].count;
    return n;
}

int h2a_msm_g1_dev(h2a_ctx* ctx, const h2a_bases* bases, size_t offset, const void* d_scalars, size_t n,
                   uint8_t out_affine[64]) {
    H2A_DEVICE(ctx);
    if (!ctx || !bases || !out_affine || (!d_scalars && n)) return H2A_ERR_INVALID;
    if (offset > bases->n || n > bases->n - offset)
        H2A_FAIL(ctx, H2A_ERR_INVALID, "msm: offset %zu + n %zu exceeds %zu bases", offset, n, bases->n);
    return h2a_msm_run(ctx, bases, offset, (const uint8_t*)d_scalars, n, out_affine);
}

// A large MSM whose scalars sit in host memory is cut into point ranges that alternate between the two lanes: while
// one range is sorted and accumulated, the next range's scalars cross PCIe.  The copies are chained (one at a time on
// the link), each range yields a canonical affine partial, and the partials are added on the host in range order.
static int msm_host_split(h2a_ctx* ctx, const h2a_bases* bases, size_t offset, const uint8_t* scalars, size_t n,
                          uint8_t out_affine[64]) {
    h2a_ctx* alt = nullptr;
    H2A_TRY(h2a_get_alt(ctx, &alt));
    alt->msm_window_override = ctx->msm_window_override;
    alt->msm_algo = ctx->msm_algo;
    h2a_ctx* lanes[2] = {ctx, alt};
    const int pieces = ctx->msm_host_split;
    // the first range is the only one whose copy is exposed: it gets a smaller share, the rest is cut evenly (two ranges, 2^22
    // points, measured per call: 35 % 10.95 ms, 42 % 10.90, 50 % 11.06, 58 % 11.45)
    int first_pct = pieces == 2 ? 42 : 100 / pieces;
    if (const char* env = getenv("H2A_MSM_HOST_FIRST_PCT")) first_pct = std::max(1, std::min(99, atoi(env)));
    const size_t first_len = std::min(n, (n * (size_t)first_pct / 100 + 1023) & ~(size_t)1023);
    const size_t piece_len = pieces > 1 ? (n - first_len + pieces - 2) / (pieces - 1) : 0;
    std::vector<uint8_t> partials((size_t)pieces * 64, 0);
    struct Events {   // released on every exit path
        cudaEvent_t e[2] = {nullptr, nullptr};
        ~Events() { for (cudaEvent_t x : e) if (x) cudaEventDestroy(x); }
        cudaEvent_t& operator[](int i) { return e[i]; }
    } copied;
    for (int l = 0; l < 2; l++) H2A_CUDA(ctx, cudaEventCreateWithFlags(&copied[l], cudaEventDisableTiming));
    H2A_CUDA(ctx, cudaEventRecord(copied[1], ctx->stream));   // the second lane starts after work already queued here
    H2A_CUDA(ctx, cudaStreamWaitEvent(alt->stream, copied[1], 0));
    int pending_piece[2] = {-1, -1};
    const bool prof = ctx->profiling;
    ctx->profiling = false;
    int rc = H2A_OK;
    for (int j = 0; j < pieces && rc == H2A_OK; j++) {
        const size_t lo = j == 0 ? 0 : std::min(n, first_len + (size_t)(j - 1) * piece_len);
        const size_t len = j == 0 ? first_len : std::min(n - lo, piece_len);
        h2a_ctx* lane = lanes[j & 1];
        if (pending_piece[j & 1] >= 0) {
            rc = h2a_msm_finish(lane, partials.data() + 64 * pending_piece[j & 1]);
            pending_piece[j & 1] = -1;
            if (rc != H2A_OK) break;
        }
        rc = h2a_reserve(lane, lane->scalars, len * 32 + 32);
        if (rc == H2A_OK && len) {
            cudaError_t e = cudaSuccess;
            if (j > 0) e = cudaStreamWaitEvent(lane->stream, copied[(j - 1) & 1], 0);   // one copy on the link at a time
            if (e == cudaSuccess) e = cudaMemcpyAsync(lane->scalars.p, scalars + 32 * lo, len * 32, cudaMemcpyHostToDevice, lane->stream);
            if (e == cudaSuccess) e = cudaEventRecord(copied[j & 1], lane->stream);
            if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = H2A_ERR_CUDA; }
        }
        if (rc == H2A_OK) rc = h2a_msm_launch(lane, bases, offset + lo, (const uint8_t*)lane->scalars.p, len);
        if (rc == H2A_OK) pending_piece[j & 1] = j;
        else if (lane != ctx) ctx->err = lane->err;
    }
    for (int l = 0; l < 2; l++) {
        if (pending_piece[l] >= 0) {
            int r2 = h2a_msm_finish(lanes[l], partials.data() + 64 * pending_piece[l]);
            if (rc == H2A_OK) rc = r2;
        }
    }
    ctx->profiling = prof;
    ctx->launches += alt->launches;
    alt->launches = 0;
    if (rc != H2A_OK) return rc;
    return h2a_g1_sum(partials.data(), (size_t)pieces, out_affine);
}

int h2a_msm_g1(h2a_ctx* ctx, const h2a_bases* bases, size_t offset, const uint8_t* scalars, size_t n,
               uint8_t out_affine[64]) {
    H2A_DEVICE(ctx);
    if (!ctx || !bases || !out_affine || (!scalars && n)) return H2A_ERR_INVALID;
    if (offset > bases->n || n > bases->n - offset)
        H2A_FAIL(ctx, H2A_ERR_INVALID, "msm: offset %zu + n %zu exceeds %zu bases", offset, n, bases->n);
    if (n == 0) {
        memset(out_affine, 0, 64);
        return H2A_OK;
    }
    if (n >= ((size_t)1 << 21) && ctx->msm_host_split > 1) return msm_host_split(ctx, bases, offset, scalars, n, out_affine);
    H2A_TRY(h2a_reserve(ctx, ctx->scalars, n * 32));
    H2A_CUDA(ctx, cudaMemcpyAsync(ctx->scalars.p, scalars, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    return h2a_msm_run(ctx, bases, offset, (const uint8_t*)ctx->scalars.p, n, out_affine);
}

int h2a_msm_g1_batch_dev(h2a_ctx* ctx, const h2a_bases* bases, const void* const* d_scalars, const size_t* n, int m,
                         uint8_t* out_affine) {
    H2A_DEVICE(ctx);
    if (!ctx || !bases || !d_scalars || !n || !out_affine || m < 0) return H2A_ERR_INVALID;
    for (int j = 0; j < m; j++)
        if (n[j] > bases->n) H2A_FAIL(ctx, H2A_ERR_INVALID, "msm batch: column %d has %zu scalars for %zu bases", j, n[j], bases->n);
    return h2a_msm_batch_dev(ctx, bases, (const uint8_t* const*)d_scalars, n, m, out_affine);
}

// host scalars: column j is copied on its lane's stream while the other lane computes
int h2a_msm_g1_batch(h2a_ctx* ctx, const h2a_bases* bases, const uint8_t* const* scalars, const size_t* n, int m,
                     uint8_t* out_affine) {
    H2A_DEVICE(ctx);
    if (!ctx || !bases || !scalars || !n || !out_affine || m < 0) return H2A_ERR_INVALID;
    for (int j = 0; j < m; j++)
        if (n[j] > bases->n) H2A_FAIL(ctx, H2A_ERR_INVALID, "msm batch: column %d has %zu scalars for %zu bases", j, n[j], bases->n);
    if (m <= 1) {
        for (int j = 0; j < m; j++) H2A_TRY(h2a_msm_g1(ctx, bases, 0, scalars[j], n[j], out_affine + 64 * j));
        return H2A_OK;
    }
    h2a_ctx* alt = nullptr;
    H2A_TRY(h2a_get_alt(ctx, &alt));
    alt->msm_window_override = ctx->msm_window_override;
    alt->msm_algo = ctx->msm_algo;
    h2a_ctx* lanes[2] = {ctx, alt};
    int pending_col[2] = {-1, -1};
    const bool prof = ctx->profiling;
    ctx->profiling = false;
    int rc = H2A_OK;
    for (int j = 0; j < m && rc == H2A_OK; j++) {
        h2a_ctx* lane = lanes[j & 1];
        if (pending_col[j & 1] >= 0) {
            rc = h2a_msm_finish(lane, out_affine + 64 * pending_col[j & 1]);
            pending_col[j & 1] = -1;
            if (rc != H2A_OK) break;
        }
        rc = h2a_reserve(lane, lane->scalars, n[j] * 32 + 32);
        if (rc == H2A_OK && n[j]) {
            cudaError_t e = cudaMemcpyAsync(lane->scalars.p, scalars[j], n[j] * 32, cudaMemcpyHostToDevice, lane->stream);
            if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = H2A_ERR_CUDA; }
        }
        if (rc == H2A_OK) rc = h2a_msm_launch(lane, bases, 0, (const uint8_t*)lane->scalars.p, n[j]);
        if (rc == H2A_OK) pending_col[j & 1] = j;
        else if (lane != ctx) ctx->err = lane->err;
    }
    for (int l = 0; l < 2; l++) {
        if (pending_col[l] >= 0) {
            int r2 = h2a_msm_finish(lanes[l], out_affine + 64 * pending_col[l]);
            if (rc == H2A_OK) rc = r2;
        }
    }
    ctx->profiling = prof;
    ctx->launches += alt->launches;
    alt->launches = 0;
    return rc;
}

int h2a_msm_g1_adhoc(h2a_ctx* ctx, const uint8_t* bases_affine, const uint8_t* scalars, size_t n, uint8_t out_affine[64]) {
    H2A_DEVICE(ctx);
    if (!ctx || !out_affine || ((!bases_affine || !scalars) && n)) return H2A_ERR_INVALID;
    if (n == 0) {
        memset(out_affine, 0, 64);
        return H2A_OK;
    }
    if (n <= SMALL_MSM_MAX) {
        uint32_t offs[2] = {0, (uint32_t)n};
        return h2a_small_msm(ctx, bases_affine, scalars, offs, 1, out_affine);
    }
    h2a_bases* b = nullptr;
    H2A_TRY(h2a_bases_upload(ctx, bases_affine, n, &b));
    int rc = h2a_msm_g1(ctx, b, 0, scalars, n, out_affine);
    h2a_bases_free(ctx, b);
    return rc;
}

int h2a_g1_sum(const uint8_t* points_affine, size_t m, uint8_t out_affine[64]) {
    if (!out_affine || (!points_affine && m)) return H2A_ERR_INVALID;
    using namespace h2a_host;
    PointX acc = px_identity();
    for (size_t i = 0; i < m; i++) acc = px_add(acc, px_from_affine(affine_load(points_affine + 64 * i)));
    affine_store(out_affine, px_to_affine(acc));
    return H2A_OK;
}

// ------------------------------------------------------------------ NTT
int h2a_fr_root_of_unity(uint32_t k, uint8_t out[32]) {
    if (!out || k > 28) return H2A_ERR_INVALID;
    h2a_host::fr_store(out, h2a_host::fr_root_of_unity((int)k));
    return H2A_OK;
}

int h2a_ntt_dev(h2a_ctx* ctx, void* d_a, uint32_t log_n, const uint8_t omega[32], int inverse, const uint8_t* coset_shift) {
    H2A_DEVICE(ctx);
    if (!ctx || !d_a || !omega) return H2A_ERR_INVALID;
    if (log_n < 1 || log_n > 28) H2A_FAIL(ctx, H2A_ERR_INVALID, "ntt: log_n=%u not in 1..28", log_n);
    const size_t bytes = 32ull << log_n;
    H2A_TRY(h2a_reserve(ctx, ctx->ntt_b, bytes));
    if (log_n >= 12) {
        // at least two passes (a pass covers at most 2^10 points): the first reads d_a and writes the workspace, the
        // last one writes the natural-order result back into d_a — no copy
        H2A_TRY(h2a_ntt_run(ctx, (const uint8_t*)d_a, 1u << log_n, (uint8_t*)ctx->ntt_b.p, (uint8_t*)d_a, log_n, omega, inverse,
                            coset_shift));
    } else {
        H2A_TRY(h2a_ntt_run(ctx, (const uint8_t*)d_a, 1u << log_n, (uint8_t*)d_a, (uint8_t*)ctx->ntt_b.p, log_n, omega, inverse,
                            coset_shift));
        H2A_CUDA(ctx, cudaMemcpyAsync(d_a, ctx->ntt_b.p, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    H2A_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return H2A_OK;
}

int h2a_ntt(h2a_ctx* ctx, uint8_t* a, uint32_t log_n, const uint8_t omega[32], int inverse, const uint8_t* coset_shift) {
    H2A_DEVICE(ctx);
    if (!ctx || !a || !omega) return H2A_ERR_INVALID;
    if (log_n < 1 || log_n > 28) H2A_FAIL(ctx, H2A_ERR_INVALID, "ntt: log_n=%u not in 1..28", log_n);
    const size_t bytes = 32ull << log_n;
    H2A_TRY(h2a_reserve(ctx, ctx->ntt_a, bytes));
    H2A_TRY(h2a_reserve(ctx, ctx->ntt_b, bytes));
    H2A_CUDA(ctx, cudaMemcpyAsync(ctx->ntt_a.p, a, bytes, cudaMemcpyHostToDevice, ctx->stream));
    H2A_TRY(h2a_ntt_run(ctx, (const uint8_t*)ctx->ntt_a.p, 1u << log_n, (uint8_t*)ctx->ntt_a.p, (uint8_t*)ctx->ntt_b.p,
                        log_n, omega, inverse, coset_shift));
    H2A_CUDA(ctx, cudaMemcpyAsync(a, ctx->ntt_b.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    H2A_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return H2A_OK;
}

int h2a_coeff_to_extended(h2a_ctx* ctx, const uint8_t* coeffs, uint32_t k, uint32_t ext_k, const uint8_t coset_shift[32],
                          uint8_t* out) {
    H2A_DEVICE(ctx);
    if (!ctx || !coeffs || !out || !coset_shift) return H2A_ERR_INVALID;
    if (k < 1 || ext_k < k || ext_k > 28) H2A_FAIL(ctx, H2A_ERR_INVALID, "coeff_to_extended: k=%u ext_k=%u", k, ext_k);
    const size_t in_bytes = 32ull << k, out_bytes = 32ull << ext_k;
    H2A_TRY(h2a_reserve(ctx, ctx->scalars, in_bytes));
    H2A_TRY(h2a_reserve(ctx, ctx->ntt_a, out_bytes));
    H2A_TRY(h2a_reserve(ctx, ctx->ntt_b, out_bytes));
    H2A_CUDA(ctx, cudaMemcpyAsync(ctx->scalars.p, coeffs, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    uint8_t omega[32];
    h2a_fr_root_of_unity(ext_k, omega);
    H2A_TRY(h2a_ntt_run(ctx, (const uint8_t*)ctx->scalars.p, 1u << k, (uint8_t*)ctx->ntt_a.p, (uint8_t*)ctx->ntt_b.p, ext_k,
                        omega, 0, coset_shift));
    H2A_CUDA(ctx, cudaMemcpyAsync(out, ctx->ntt_b.p, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    H2A_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return H2A_OK;
}

int h2a_coeff_to_extended_dev(h2a_ctx* ctx, const void* d_coeffs, uint32_t k, uint32_t ext_k, const uint8_t coset_shift[32],
                              void* d_out) {
    H2A_DEVICE(ctx);
    if (!ctx || !d_coeffs || !d_out || !coset_shift) return H2A_ERR_INVALID;
    if (k < 1 || ext_k < k || ext_k > 28) H2A_FAIL(ctx, H2A_ERR_INVALID, "coeff_to_extended: k=%u ext_k=%u", k, ext_k);
    const size_t in_bytes = 32ull << k, out_bytes = 32ull << ext_k;
    const uint8_t *lo = (const uint8_t*)d_coeffs, *out = (const uint8_t*)d_out;
    if (lo < out + out_bytes && out < lo + in_bytes) H2A_FAIL(ctx, H2A_ERR_INVALID, "coeff_to_extended: input and output overlap");
    H2A_TRY(h2a_reserve(ctx, ctx->ntt_a, out_bytes));
    uint8_t omega[32];
    h2a_fr_root_of_unity(ext_k, omega);
    H2A_TRY(h2a_ntt_run(ctx, lo, 1u << k, (uint8_t*)ctx->ntt_a.p, (uint8_t*)d_out, ext_k, omega, 0, coset_shift));
    H2A_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return H2A_OK;
}

int h2a_extended_to_coeff_dev(h2a_ctx* ctx, void* d_ext, uint32_t ext_k, const uint8_t coset_shift[32]) {
    H2A_DEVICE(ctx);
    if (!ctx || !d_ext || !coset_shift) return H2A_ERR_INVALID;
    uint8_t omega[32];
    if (ext_k < 1 || h2a_fr_root_of_unity(ext_k, omega) != H2A_OK) H2A_FAIL(ctx, H2A_ERR_INVALID, "extended_to_coeff: ext_k=%u", ext_k);
    return h2a_ntt_dev(ctx, d_ext, ext_k, omega, 1, coset_shift);
}

int h2a_extended_to_coeff(h2a_ctx* ctx, uint8_t* ext, uint32_t ext_k, const uint8_t coset_shift[32]) {
    H2A_DEVICE(ctx);
    if (!ctx || !ext || !coset_shift) return H2A_ERR_INVALID;
    uint8_t omega[32];
    if (h2a_fr_root_of_unity(ext_k, omega) != H2A_OK || ext_k < 1)
        H2A_FAIL(ctx, H2A_ERR_INVALID, "extended_to_coeff: ext_k=%u", ext_k);
    return h2a_ntt(ctx, ext, ext_k, omega, 1, coset_shift);
}

}  // extern "C"
