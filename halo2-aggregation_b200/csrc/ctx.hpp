// ctx.hpp — the library context behind the opaque `h2a_ctx` of include/h2agg.h.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "../../include/h2agg.h"

// A growable device buffer owned by the ctx (workspace that persists across calls so the hot path
// does no cudaMalloc after warm-up).
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct NttTables;  // ntt.cu

struct h2a_ctx {
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;       // second stream of this lane (the two halves of the affine tree)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaStream_t stream_lo[2] = {nullptr, nullptr};   // one priority level below: the tree's backward passes (msm_tree_prio)
    cudaEvent_t ev_tree[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    int msm_tree_prio = 0;                // 1: backward passes of the addition tree on lower-priority streams (H2A_MSM_TREE_PRIO)
    int stream_priority = 0;              // priority both streams are created with (the device's greatest)
    std::string err;
    uint64_t launches = 0;
    int msm_window_override = 0;
    bool ntt_attr_set = false, sort_attr_set = false;  // cudaFuncSetAttribute done for this ctx's device
    int ntt_log_tile = 10;  // log2 of the elements one NTT block holds in shared memory (32 B each)
    int msm_red_chunk = 2048;  // segments per block at level 2 of the bucket reduction (power of two)
    int msm_seg_len = 16;  // buckets per thread at level 1 of the bucket reduction (power of two)
    int msm_host_split = 2;  // point ranges a large host-scalar MSM is cut into so copies overlap compute (1 = off)
    int msm_group_cols = 8;       // columns one pass of a device-resident batch takes (tables only; 1 = one MSM per column)
    int msm_group_cols_host = 2;  // the same when the columns are copied from host memory on the way
    int msm_rounds_bias = 0;      // experiment knob (H2A_MSM_ROUNDS_BIAS): added to the number of regular tree rounds
    int msm_algo = 1;  // 0: XYZZ mixed additions, one thread per bucket task; 1: pairwise tree of batched affine additions

    // profiling
    bool profiling = false;
    int last_kind = 0;  // 0 msm, 1 ntt
    std::vector<cudaEvent_t> ev;
    int ev_used = 0;
    std::vector<float> phase_ms;
    std::vector<const char*> prove_phase_names;

    // MSM workspace
    DevBuf scalars, offsets, cursor, sorted, buckets, segsums, winsums, heavy, misc, aff_a, aff_b, aff_c, aff_scratch, aff_u32;
    void* pinned = nullptr;  // small pinned staging area for results
    size_t pinned_cap = 0;

    // an MSM whose kernels are queued but whose window sums have not been combined yet (h2a_msm_launch / _finish)
    struct {
        bool active = false;
        bool pre = false;
        int c = 0;
        uint32_t groups = 0;   // points per column waiting in the pinned buffer
        uint32_t cols = 1;     // columns of this pass (h2a_msm_launch_cols)
        uint32_t log_l = 0;  // log2 of the reduction's segment length (the host finishes the single-window Horner)
    } msm_pending;
    h2a_ctx* alt = nullptr;  // second lane (own stream + workspace) for pipelined batches; created on first use

    // multi-GPU plumbing (comm.cu): two NCCL communicators (0: small exchanges on the ctx stream, 1: bulk broadcasts on the
    // prover's transform lane), this process's rank, and a staging buffer for host-buffer collectives
    void* comm[2] = {nullptr, nullptr};
    int comm_rank = 0, comm_world = 1;
    uint64_t comm_bytes[2] = {0, 0};      // payload bytes this rank has put into / taken out of collectives since comm_init
    DevBuf comm_buf;

    // NTT workspace
    DevBuf ntt_a, ntt_b;
    std::map<uint32_t, NttTables*> ntt_tables;  // keyed by log_n of the twiddle table
};

struct h2a_bases {
    const uint8_t* d = nullptr;  // n * 64 bytes, device
    size_t n = 0;
    bool owned = false;
    uint8_t* table = nullptr;  // optional precomputed window tables T[w][i] = 2^(c*w) P_i, W * n * 64 bytes
    int table_c = 0;
};

#define H2A_FAIL(ctx, code, ...)                          \
    do {                                                  \
        char _b[512];                                     \
        snprintf(_b, sizeof _b, __VA_ARGS__);             \
        (ctx)->err = _b;                                  \
        return (code);                                    \
    } while (0)

#define H2A_CUDA(ctx, call)                                                                              \
    do {                                                                                                 \
        cudaError_t _e = (call);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            char _b[512];                                                                                \
            snprintf(_b, sizeof _b, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            (ctx)->err = _b;                                                                             \
            return _e == cudaErrorMemoryAllocation ? H2A_ERR_OOM : H2A_ERR_CUDA;                         \
        }                                                                                                \
    } while (0)

// Every public entry point first binds the calling thread to the ctx's GPU: two ctxs on different GPUs in one process, or a host
// framework that changed the current device, must not send allocations and launches to the wrong device.
#define H2A_DEVICE(ctx)                           \
    do {                                          \
        if (ctx) cudaSetDevice((ctx)->device);    \
    } while (0)

#define H2A_TRY(expr)             \
    do {                          \
        int _rc = (expr);         \
        if (_rc != H2A_OK) return _rc; \
    } while (0)

// grow-only reservation
int h2a_reserve(h2a_ctx* ctx, DevBuf& b, size_t bytes);
int h2a_reserve_pinned(h2a_ctx* ctx, size_t bytes);

// profiling helpers: phase_begin() resets, phase_mark() records an event after a phase
void h2a_prof_begin(h2a_ctx* ctx, int kind);
void h2a_prof_mark(h2a_ctx* ctx);
void h2a_prof_end(h2a_ctx* ctx);

#define H2A_LAUNCH_CHECK(ctx)                  \
    do {                                       \
        (ctx)->launches++;                     \
        H2A_CUDA(ctx, cudaGetLastError());     \
    } while (0)

// implemented in msm.cu / ntt.cu / misc.cu
int h2a_msm_run(h2a_ctx* ctx, const h2a_bases* bases, size_t offset, const uint8_t* d_scalars, size_t n,
                uint8_t out_affine[64]);
int h2a_msm_precompute(h2a_ctx* ctx, h2a_bases* bases, int c);
// the two halves of h2a_msm_run: queue every kernel and the copy of the window sums (no host synchronisation), then
// wait and combine on the host.  One MSM may be pending per ctx (lane).
int h2a_msm_launch(h2a_ctx* ctx, const h2a_bases* bases, size_t offset, const uint8_t* d_scalars, size_t n);
int h2a_msm_launch_cols(h2a_ctx* ctx, const h2a_bases* bases, const uint8_t* const* d_cols, int m, size_t n);
int h2a_msm_finish(h2a_ctx* ctx, uint8_t* out_affine /* 64 bytes per column of the pending pass */);
// m MSMs over the same bases with device-resident scalars, pipelined over two lanes
int h2a_msm_batch_dev(h2a_ctx* ctx, const h2a_bases* bases, const uint8_t* const* d_scalars, const size_t* n, int m, uint8_t* out_affine,
                      const uint8_t* const* h_src = nullptr);
int h2a_get_alt(h2a_ctx* ctx, h2a_ctx** out);
// comm.cu
bool h2a_comm_active(const h2a_ctx* ctx);
int h2a_comm_group_start(h2a_ctx* ctx);
int h2a_comm_group_end(h2a_ctx* ctx);
int h2a_comm_broadcast_on(h2a_ctx* ctx, int lane, void* d_buf, size_t bytes, int root, cudaStream_t stream);
int h2a_comm_send_on(h2a_ctx* ctx, int lane, const void* d_buf, size_t bytes, int peer, cudaStream_t stream);
int h2a_comm_recv_on(h2a_ctx* ctx, int lane, void* d_buf, size_t bytes, int peer, cudaStream_t stream);
int h2a_comm_allgather_on(h2a_ctx* ctx, int lane, const void* d_send, void* d_recv, size_t bytes_per_rank, cudaStream_t stream);
