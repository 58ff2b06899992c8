// Tile staging for a streaming field kernel: plain 128-bit loads against 1-D bulk copies (TMA, cp.async.bulk + mbarrier).
// Workload: an array of 2^24 Fr elements (512 MB, larger than L2); every element is multiplied MULS times by a constant
// (MULS = 4 is one NTT pass: ~3 butterfly products + one twist per element; 0 is a pure copy) and written back.
//   ldg   each thread loads its elements with 2 x LDG.128, computes, stores
//   bulk  one elected thread per block issues a 32 KB cp.async.bulk for tile i+1 into the other half of a double buffer
//         while the block works on tile i out of shared memory (mbarrier complete_tx signalling); results stored with STG
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I../../halo2-aggregation_b200/csrc -o _bin/bulk_copy bulk_copy.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "field.cuh"
using namespace h2a;

constexpr int TILE = 1024;          // elements per tile (32 KB)
constexpr int THREADS = 256;

template <int MULS>
__device__ __forceinline__ Fr work(Fr v, const Fr& c) {
#pragma unroll
    for (int k = 0; k < MULS; k++) v = v * c;
    return v;
}

template <int MULS>
__global__ void __launch_bounds__(THREADS) ldg_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, uint32_t tiles) {
    const Fr c = Fr::r2();
    for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const size_t base = (size_t)tile * TILE;
#pragma unroll
        for (int k = 0; k < TILE / THREADS; k++) {
            const size_t i = base + k * THREADS + threadIdx.x;
            work<MULS>(Fr::load(in + 32 * i), c).store(out + 32 * i);
        }
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MULS>
__global__ void __launch_bounds__(THREADS) bulk_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, uint32_t tiles) {
    extern __shared__ __align__(128) uint8_t smem[];          // 2 x 32 KB tiles
    __shared__ __align__(8) uint64_t bar[2];
    const Fr c = Fr::r2();
    const uint32_t bytes = TILE * 32;
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; b++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[b])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](uint32_t tile, int b) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[b])), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + (size_t)b * bytes)),
                     "l"(in + (size_t)tile * bytes), "r"(bytes), "r"(smem_u32(&bar[b]))
                     : "memory");
    };
    if (threadIdx.x == 0 && blockIdx.x < tiles) issue(blockIdx.x, 0);
    uint32_t it = 0;
    for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, it++) {
        const int b = it & 1;
        const uint32_t next = tile + gridDim.x;
        if (threadIdx.x == 0 && next < tiles) issue(next, b ^ 1);     // the other half was released by the barrier below
        const uint32_t parity = (it >> 1) & 1;
        uint32_t ok;
        do {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bar[b])), "r"(parity) : "memory");
        } while (!ok);
        const uint8_t* src = smem + (size_t)b * bytes;
        const size_t base = (size_t)tile * TILE;
#pragma unroll
        for (int k = 0; k < TILE / THREADS; k++) {
            const uint32_t i = k * THREADS + threadIdx.x;
            work<MULS>(Fr::load(src + 32 * i), c).store(out + 32 * (base + i));
        }
        __syncthreads();                                              // everyone is done with this half before it is refilled
    }
}

template <int MULS>
static void run(const uint8_t* in, uint8_t* out, uint32_t n) {
    const uint32_t tiles = n / TILE;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float ms_l = 0, ms_b = 0;
    const int grid = 148 * 3;
    cudaFuncSetAttribute(bulk_kernel<MULS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE * 32);
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(a);
        ldg_kernel<MULS><<<grid, THREADS>>>(in, out, tiles);
        cudaEventRecord(b); cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms_l, a, b);
        cudaEventRecord(a);
        bulk_kernel<MULS><<<grid, THREADS, 2 * TILE * 32>>>(in, out, tiles);
        cudaEventRecord(b); cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms_b, a, b);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("{\"products_per_element\": %d, \"ldg_ms\": %.3f, \"ldg_gb_per_s\": %.0f, \"bulk_ms\": %.3f, \"bulk_gb_per_s\": %.0f, \"status\": \"%s\"}\n", MULS, ms_l,
           64.0 * n / (ms_l * 1e-3) / 1e9, ms_b, 64.0 * n / (ms_b * 1e-3) / 1e9, cudaGetErrorString(e));
}
int main() {
    const uint32_t n = 1u << 24;
    uint8_t *in, *out;
    if (cudaMalloc(&in, 32ull * n) != cudaSuccess || cudaMalloc(&out, 32ull * n) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(in, 1, 32ull * n);
    run<0>(in, out, n);
    run<1>(in, out, n);
    run<4>(in, out, n);
    return 0;
}
