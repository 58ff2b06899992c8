// One radix-2 butterfly level between elements held by DIFFERENT lanes, 256-bit (8-word) field elements:
//   shfl  partner's element fetched with 8 x SHFL.BFLY (one per word), no shared memory, no barrier
//   smem  every lane stores its element (2 x STS.128), __syncwarp, loads the partner's (2 x LDS.128)
// then u + v / (u - v) * w as in an NTT stage.  32 levels per iteration, distances 1..16 cycling.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I../../halo2-aggregation_b200/csrc -o _bin/exchange256 exchange256.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "field.cuh"
using namespace h2a;

template <bool SHFL, bool MUL>
__global__ void __launch_bounds__(256) level_kernel(int iters, uint8_t* out) {
    __shared__ uint4 lo[256], hi[256];
    Fr v = Fr::one();
    v.l[0] ^= threadIdx.x + blockIdx.x * 977u;
    const Fr w = Fr::r2();
    const uint32_t lane = threadIdx.x & 31u;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int lvl = 0; lvl < 5; lvl++) {
            const uint32_t d = 1u << lvl;
            Fr p;
            if (SHFL) {
#pragma unroll
                for (int k = 0; k < 8; k++) p.l[k] = __shfl_xor_sync(0xffffffffu, v.l[k], d);
            } else {
                lo[threadIdx.x] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
                hi[threadIdx.x] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
                __syncwarp();
                const uint4 a = lo[threadIdx.x ^ d], b = hi[threadIdx.x ^ d];
                p.l[0] = a.x; p.l[1] = a.y; p.l[2] = a.z; p.l[3] = a.w; p.l[4] = b.x; p.l[5] = b.y; p.l[6] = b.z; p.l[7] = b.w;
                __syncwarp();
            }
            if (lane & d) { Fr t = p - v; v = MUL ? t * w : t; }
            else v = v + p;
        }
    }
    v.store(out + 32ull * (blockIdx.x * blockDim.x + threadIdx.x));
}
template <bool SHFL, bool MUL>
static double run(uint8_t* out) {
    const int blocks = 148 * 8, iters = 400;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    level_kernel<SHFL, MUL><<<blocks, 256>>>(iters, out);
    cudaEventRecord(a);
    level_kernel<SHFL, MUL><<<blocks, 256>>>(iters, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return 5.0 * iters * blocks * 256 / (ms * 1e-3) / 1e9;     // element-levels per second (1e9)
}
int main() {
    uint8_t* out;
    cudaMalloc(&out, 32ull * 148 * 8 * 256);
    printf("{\"exchange_only\": {\"shfl_gelem_levels_per_s\": %.1f, \"smem_gelem_levels_per_s\": %.1f},", run<true, false>(out), run<false, false>(out));
    printf(" \"with_twiddle_product\": {\"shfl_gelem_levels_per_s\": %.1f, \"smem_gelem_levels_per_s\": %.1f}}\n", run<true, true>(out), run<false, true>(out));
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
