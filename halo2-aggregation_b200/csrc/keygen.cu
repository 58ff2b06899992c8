// keygen.cu — the permutation part of key generation: copy constraints -> sigma columns.
//
// `keygen_vk` / `keygen_pk` (examples/simple-example.rs:593-594, :696-697) turn the copy constraints recorded while
// the circuit is laid out into one permutation of the cells of the permutation columns
// (`vk.permutation`, read back at src/verifier.rs:244-259 and src/permutation.rs:259 with `Fr::DELTA`).  The
// sigma column j holds, at row i, the label delta^{j'} * omega^{i'} of the cell (j', i') that follows (j, i) in
// its cycle.  The cycle bookkeeping is sequential host work; the n_cols * n labels are one element-wise kernel.
// The sigma columns produced here are what h2a_circuit_set_keys takes.
#include <cstring>
#include <utility>
#include <vector>

#include "ctx.hpp"
#include "field.cuh"
#include "host_bn254.hpp"

int h2a_pow_vector(h2a_ctx* ctx, const uint8_t base[32], const uint8_t c[32], uint32_t n, uint8_t* d_out);

// Cells are numbered col * n + row.  `mapping` is the permutation (next cell of the cycle), `aux` names each
// cell's cycle by one of its members, `sizes` is the length of the cycle a name stands for.
struct h2a_assembly {
    uint32_t n_cols = 0, k = 0;
    std::vector<uint32_t> mapping, aux, sizes;
};

namespace dev {
using namespace h2a;

__global__ void sigma_labels_kernel(const uint32_t* __restrict__ mapping, uint32_t cells, uint32_t k,
                                    const uint8_t* __restrict__ omega_pows, const uint8_t* __restrict__ delta_pows,
                                    uint8_t* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cells) return;
    const uint32_t m = mapping[i];
    const Fp<FR> w = Fp<FR>::load(omega_pows + 32ull * (m & ((1u << k) - 1)));
    const Fp<FR> d = Fp<FR>::load(delta_pows + 32ull * (m >> k));
    (w * d).store(out + 32ull * i);
}
}  // namespace dev

int h2a_assembly_new(uint32_t n_cols, uint32_t k, h2a_assembly** out) {
    if (!out || k > 27 || n_cols == 0 || ((uint64_t)n_cols << k) >= (1ull << 32)) return H2A_ERR_INVALID;
    h2a_assembly* a = new h2a_assembly;
    a->n_cols = n_cols;
    a->k = k;
    const size_t cells = (size_t)n_cols << k;
    a->mapping.resize(cells);
    a->aux.resize(cells);
    a->sizes.assign(cells, 1);
    for (size_t i = 0; i < cells; i++) a->mapping[i] = a->aux[i] = (uint32_t)i;
    *out = a;
    return H2A_OK;
}

void h2a_assembly_free(h2a_assembly* a) { delete a; }

// Joins the cycles of the two cells: the shorter cycle is renamed into the longer one, then exchanging the two
// successors splices the cycles together.  A no-op when both cells already share a cycle.
int h2a_assembly_copy(h2a_assembly* a, uint32_t col_a, uint32_t row_a, uint32_t col_b, uint32_t row_b) {
    if (!a) return H2A_ERR_INVALID;
    const uint32_t n = 1u << a->k;
    if (col_a >= a->n_cols || col_b >= a->n_cols || row_a >= n || row_b >= n) return H2A_ERR_INVALID;
    uint32_t left = col_a * n + row_a, right = col_b * n + row_b;
    if (a->aux[left] == a->aux[right]) return H2A_OK;
    if (a->sizes[a->aux[left]] < a->sizes[a->aux[right]]) std::swap(left, right);
    const uint32_t name = a->aux[left];
    a->sizes[name] += a->sizes[a->aux[right]];
    for (uint32_t cell = right; a->aux[cell] != name; cell = a->mapping[cell]) a->aux[cell] = name;
    std::swap(a->mapping[left], a->mapping[right]);
    return H2A_OK;
}

int h2a_assembly_mapping(const h2a_assembly* a, uint32_t* out_next_cell) {
    if (!a || !out_next_cell) return H2A_ERR_INVALID;
    memcpy(out_next_cell, a->mapping.data(), a->mapping.size() * 4);
    return H2A_OK;
}

namespace {
struct Scoped {
    void* p = nullptr;
    ~Scoped() { if (p) cudaFree(p); }
};
}  // namespace

int h2a_assembly_sigmas(h2a_ctx* ctx, const h2a_assembly* a, const uint8_t omega[32], const uint8_t delta[32],
                        uint8_t* out_sigmas) {
    H2A_DEVICE(ctx);
    if (!ctx || !a || !omega || !delta || !out_sigmas) return H2A_ERR_INVALID;
    namespace hh = h2a_host;
    const uint32_t n = 1u << a->k;
    const size_t cells = a->mapping.size();
    std::vector<uint8_t> dpow(32 * (size_t)a->n_cols);
    hh::Fr acc = hh::fr_one();
    const hh::Fr d = hh::fr_load(delta);
    for (uint32_t j = 0; j < a->n_cols; j++) {
        hh::fr_store(dpow.data() + 32 * (size_t)j, acc);
        acc = acc * d;
    }
    uint8_t one[32];
    hh::fr_store(one, hh::fr_one());
    Scoped d_map, d_w, d_d, d_out;
    H2A_CUDA(ctx, cudaMalloc(&d_map.p, cells * 4));
    H2A_CUDA(ctx, cudaMalloc(&d_w.p, (size_t)n * 32));
    H2A_CUDA(ctx, cudaMalloc(&d_d.p, dpow.size()));
    H2A_CUDA(ctx, cudaMalloc(&d_out.p, cells * 32));
    H2A_CUDA(ctx, cudaMemcpyAsync(d_map.p, a->mapping.data(), cells * 4, cudaMemcpyHostToDevice, ctx->stream));
    H2A_CUDA(ctx, cudaMemcpyAsync(d_d.p, dpow.data(), dpow.size(), cudaMemcpyHostToDevice, ctx->stream));
    H2A_TRY(h2a_pow_vector(ctx, omega, one, n, (uint8_t*)d_w.p));
    dev::sigma_labels_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, ctx->stream>>>(
        (const uint32_t*)d_map.p, (uint32_t)cells, a->k, (const uint8_t*)d_w.p, (const uint8_t*)d_d.p, (uint8_t*)d_out.p);
    H2A_LAUNCH_CHECK(ctx);
    H2A_CUDA(ctx, cudaMemcpyAsync(out_sigmas, d_out.p, cells * 32, cudaMemcpyDeviceToHost, ctx->stream));
    H2A_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return H2A_OK;
}
