// comm.cu — the library's own multi-GPU plumbing: one process per GPU, one ctx per process, NCCL over
// NVLink / NVSwitch for the few exchanges the path has — 64-byte partial MSM results and commitments (allgather of raw
// bytes), and, for one proof over several GPUs, whole polynomials broadcast from the rank that transformed them.
//
// A Rust host needs nothing but this library: rank 0 asks for a unique id (h2a_comm_unique_id), hands the 128 bytes to
// the other processes over whatever channel it has (a file, a socket, MPI), and every process calls h2a_comm_init.
// NCCL is bound at run time (dlopen of libnccl.so.2 — the copy a host framework such as torch has already loaded, if
// any), so the library loads on a box without it and single-GPU users pay nothing.
//
// Two communicators per ctx: lane 0 carries the small exchanges issued from the ctx stream, lane 1 the bulk broadcasts
// on the prover's transform lane.  NCCL serialises the operations of ONE communicator in issue order whatever their
// streams, so a 64-byte commitment exchange would otherwise queue behind hundreds of MB of polynomials.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "ctx.hpp"

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    const char* (*GetLastError)(ncclComm_t) = nullptr;
    bool ok = false;
};

NcclApi& nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {   // a copy already in the process first (torch ships its own)
        api.lib = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib)
        for (const char* nm : names) {
            api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
    if (!api.lib) return api;
#define BIND(field, sym) *(void**)(&api.field) = dlsym(api.lib, sym)
    BIND(GetUniqueId, "ncclGetUniqueId");
    BIND(CommInitRank, "ncclCommInitRank");
    BIND(CommDestroy, "ncclCommDestroy");
    BIND(AllGather, "ncclAllGather");
    BIND(Broadcast, "ncclBroadcast");
    BIND(Send, "ncclSend");
    BIND(Recv, "ncclRecv");
    BIND(GroupStart, "ncclGroupStart");
    BIND(GroupEnd, "ncclGroupEnd");
    BIND(GetErrorString, "ncclGetErrorString");
    BIND(GetLastError, "ncclGetLastError");
#undef BIND
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.Broadcast && api.Send && api.Recv && api.GroupStart && api.GroupEnd &&
             api.GetErrorString;
    return api;
}

}  // namespace

#define H2A_NCCL(ctx, call)                                                                           \
    do {                                                                                              \
        ncclResult_t _r = (call);                                                                     \
        if (_r != ncclSuccess) H2A_FAIL(ctx, H2A_ERR_CUDA, "%s -> %s", #call, nccl().GetErrorString(_r)); \
    } while (0)

// ---- used by the prover (plonk_prove.cu)
bool h2a_comm_active(const h2a_ctx* ctx) { return ctx && ctx->comm_world > 1 && ctx->comm[0]; }

int h2a_comm_group_start(h2a_ctx* ctx) {
    H2A_NCCL(ctx, nccl().GroupStart());
    return H2A_OK;
}
int h2a_comm_group_end(h2a_ctx* ctx) {
    H2A_NCCL(ctx, nccl().GroupEnd());
    return H2A_OK;
}
// in-place broadcast of `bytes` at d_buf from rank `root`, on `stream`, over communicator lane `lane`
int h2a_comm_broadcast_on(h2a_ctx* ctx, int lane, void* d_buf, size_t bytes, int root, cudaStream_t stream) {
    if (!h2a_comm_active(ctx)) H2A_FAIL(ctx, H2A_ERR_INVALID, "broadcast: no communicator (h2a_comm_init)");
    H2A_NCCL(ctx, nccl().Broadcast(d_buf, d_buf, bytes, ncclUint8, root, (ncclComm_t)ctx->comm[lane], stream));
    ctx->comm_bytes[root == ctx->comm_rank ? 0 : 1] += bytes;
    return H2A_OK;
}
// point-to-point pieces of an exchange (inside h2a_comm_group_start / _end; both sides enumerate their pieces in the same order)
int h2a_comm_send_on(h2a_ctx* ctx, int lane, const void* d_buf, size_t bytes, int peer, cudaStream_t stream) {
    H2A_NCCL(ctx, nccl().Send(d_buf, bytes, ncclUint8, peer, (ncclComm_t)ctx->comm[lane], stream));
    ctx->comm_bytes[0] += bytes;
    return H2A_OK;
}
int h2a_comm_recv_on(h2a_ctx* ctx, int lane, void* d_buf, size_t bytes, int peer, cudaStream_t stream) {
    H2A_NCCL(ctx, nccl().Recv(d_buf, bytes, ncclUint8, peer, (ncclComm_t)ctx->comm[lane], stream));
    ctx->comm_bytes[1] += bytes;
    return H2A_OK;
}
int h2a_comm_allgather_on(h2a_ctx* ctx, int lane, const void* d_send, void* d_recv, size_t bytes_per_rank, cudaStream_t stream) {
    if (!h2a_comm_active(ctx)) H2A_FAIL(ctx, H2A_ERR_INVALID, "allgather: no communicator (h2a_comm_init)");
    H2A_NCCL(ctx, nccl().AllGather(d_send, d_recv, bytes_per_rank, ncclUint8, (ncclComm_t)ctx->comm[lane], stream));
    ctx->comm_bytes[0] += bytes_per_rank;
    ctx->comm_bytes[1] += bytes_per_rank * (size_t)(ctx->comm_world - 1);
    return H2A_OK;
}

extern "C" {

int h2a_comm_unique_id(uint8_t out_id[128]) {
    if (!out_id) return H2A_ERR_INVALID;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    if (!nccl().ok) return H2A_ERR_NO_DEVICE;
    ncclUniqueId id;
    if (nccl().GetUniqueId(&id) != ncclSuccess) return H2A_ERR_CUDA;
    memcpy(out_id, &id, 128);
    return H2A_OK;
}

int h2a_comm_init(h2a_ctx* ctx, int rank, int world, const uint8_t id[128], const uint8_t id_bulk[128]) {
    H2A_DEVICE(ctx);
    if (!ctx || !id || !id_bulk || world < 1 || rank < 0 || rank >= world) return H2A_ERR_INVALID;
    if (ctx->comm[0]) H2A_FAIL(ctx, H2A_ERR_INVALID, "comm_init: this ctx already has a communicator");
    if (!nccl().ok) H2A_FAIL(ctx, H2A_ERR_NO_DEVICE, "comm_init: libnccl.so.2 not found (or too old)");
    if (!memcmp(id, id_bulk, 128)) H2A_FAIL(ctx, H2A_ERR_INVALID, "comm_init: the two unique ids must differ (one per communicator)");
    const uint8_t* ids[2] = {id, id_bulk};
    for (int lane = 0; lane < 2; lane++) {
        ncclUniqueId uid;
        memcpy(&uid, ids[lane], 128);
        ncclComm_t comm = nullptr;
        const ncclResult_t r = nccl().CommInitRank(&comm, world, uid, rank);
        if (r != ncclSuccess) {   // never leave a ctx with one communicator of the two
            if (ctx->comm[0]) { nccl().CommDestroy((ncclComm_t)ctx->comm[0]); ctx->comm[0] = nullptr; }
            H2A_FAIL(ctx, H2A_ERR_CUDA, "comm_init: ncclCommInitRank (communicator %d) -> %s", lane, nccl().GetErrorString(r));
        }
        ctx->comm[lane] = comm;
    }
    ctx->comm_rank = rank;
    ctx->comm_world = world;
    return H2A_OK;
}

int h2a_comm_destroy(h2a_ctx* ctx) {
    H2A_DEVICE(ctx);
    if (!ctx) return H2A_ERR_INVALID;
    for (int lane = 0; lane < 2; lane++)
        if (ctx->comm[lane]) {
            cudaStreamSynchronize(ctx->stream);
            nccl().CommDestroy((ncclComm_t)ctx->comm[lane]);
            ctx->comm[lane] = nullptr;
        }
    ctx->comm_rank = 0;
    ctx->comm_world = 1;
    return H2A_OK;
}

int h2a_comm_traffic(const h2a_ctx* ctx, uint64_t out_sent_received[2]) {
    if (!ctx || !out_sent_received) return H2A_ERR_INVALID;
    out_sent_received[0] = ctx->comm_bytes[0];
    out_sent_received[1] = ctx->comm_bytes[1];
    return H2A_OK;
}
int h2a_comm_rank(const h2a_ctx* ctx) { return ctx ? ctx->comm_rank : 0; }
int h2a_comm_world(const h2a_ctx* ctx) { return ctx ? ctx->comm_world : 1; }

int h2a_comm_allgather_dev(h2a_ctx* ctx, const void* d_send, void* d_recv, size_t bytes_per_rank) {
    H2A_DEVICE(ctx);
    if (!ctx || !d_send || !d_recv) return H2A_ERR_INVALID;
    if (ctx->comm_world == 1) {
        if (d_send != d_recv) H2A_CUDA(ctx, cudaMemcpyAsync(d_recv, d_send, bytes_per_rank, cudaMemcpyDeviceToDevice, ctx->stream));
        return H2A_OK;
    }
    return h2a_comm_allgather_on(ctx, 0, d_send, d_recv, bytes_per_rank, ctx->stream);
}

int h2a_comm_allgather(h2a_ctx* ctx, const uint8_t* send, uint8_t* recv, size_t bytes_per_rank) {
    H2A_DEVICE(ctx);
    if (!ctx || !send || !recv) return H2A_ERR_INVALID;
    const size_t world = (size_t)ctx->comm_world;
    if (world == 1) {
        memmove(recv, send, bytes_per_rank);
        return H2A_OK;
    }
    H2A_TRY(h2a_reserve(ctx, ctx->comm_buf, bytes_per_rank * (world + 1)));
    uint8_t* d = (uint8_t*)ctx->comm_buf.p;
    H2A_CUDA(ctx, cudaMemcpyAsync(d, send, bytes_per_rank, cudaMemcpyHostToDevice, ctx->stream));
    H2A_TRY(h2a_comm_allgather_on(ctx, 0, d, d + bytes_per_rank, bytes_per_rank, ctx->stream));
    H2A_CUDA(ctx, cudaMemcpyAsync(recv, d + bytes_per_rank, bytes_per_rank * world, cudaMemcpyDeviceToHost, ctx->stream));
    H2A_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return H2A_OK;
}

int h2a_comm_broadcast_dev(h2a_ctx* ctx, void* d_buf, size_t bytes, int root) {
    H2A_DEVICE(ctx);
    if (!ctx || !d_buf || root < 0 || root >= ctx->comm_world) return H2A_ERR_INVALID;
    if (ctx->comm_world == 1) return H2A_OK;
    return h2a_comm_broadcast_on(ctx, 0, d_buf, bytes, root, ctx->stream);
}

// One MSM over point ranges held by the ranks: this rank's partial over its own bases / scalars, the 64-byte affine
// partials allgathered as raw bytes, summed in rank order on every rank -> the same 64 bytes everywhere.
int h2a_msm_g1_sharded(h2a_ctx* ctx, const h2a_bases* local_bases, size_t offset, const void* d_local_scalars, size_t n_local,
                       uint8_t out_affine[64]) {
    H2A_DEVICE(ctx);
    if (!ctx || !local_bases || !out_affine || (!d_local_scalars && n_local)) return H2A_ERR_INVALID;
    if (offset > local_bases->n || n_local > local_bases->n - offset)
        H2A_FAIL(ctx, H2A_ERR_INVALID, "msm_g1_sharded: offset %zu + n %zu exceeds %zu local bases", offset, n_local, local_bases->n);
    if (ctx->comm_world > 64) H2A_FAIL(ctx, H2A_ERR_INVALID, "msm_g1_sharded: more than 64 ranks");
    uint8_t mine[64];
    H2A_TRY(h2a_msm_run(ctx, local_bases, offset, (const uint8_t*)d_local_scalars, n_local, mine));
    if (ctx->comm_world == 1) {
        memcpy(out_affine, mine, 64);
        return H2A_OK;
    }
    uint8_t all[64 * 64];
    H2A_TRY(h2a_comm_allgather(ctx, mine, all, 64));
    return h2a_g1_sum(all, (size_t)ctx->comm_world, out_affine);
}

}  // extern "C"
