// Pipe-rate micro-benchmarks for sm_100a: IMAD, IMAD.WIDE(.X), DFMA, and IMAD.WIDE + DFMA from different warps.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) imad_k(uint32_t* out, int iters, uint32_t a, uint32_t b) {
    uint32_t x[8];
    for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k;
    for (int i = 0; i < iters; i++)
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int k = 0; k < 8; k++) x[k] = x[k] * a + b;
    uint32_t s = 0;
    for (int k = 0; k < 8; k++) s ^= x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__device__ __forceinline__ void wide_body(uint64_t (&x)[8], uint32_t a) {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
        for (int k = 0; k < 8; k++) x[k] = (uint64_t)(uint32_t)x[k] * a + x[k];   // IMAD.WIDE.U32 with 64-bit accumulate
}
__global__ void __launch_bounds__(256) wide_k(uint64_t* out, int iters, uint32_t a) {
    uint64_t x[8];
    for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k;
    for (int i = 0; i < iters; i++) wide_body(x, a);
    uint64_t s = 0;
    for (int k = 0; k < 8; k++) s ^= x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__device__ __forceinline__ void dfma_body(double (&x)[8], double a, double b) {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
        for (int k = 0; k < 8; k++) x[k] = __fma_rz(x[k], a, b);
}
__global__ void __launch_bounds__(256) dfma_k(double* out, int iters, double a, double b) {
    double x[8];
    for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k;
    for (int i = 0; i < iters; i++) dfma_body(x, a, b);
    double s = 0;
    for (int k = 0; k < 8; k++) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// even warps run IMAD.WIDE, odd warps run DFMA
__global__ void __launch_bounds__(256) mixed_k(uint64_t* out, int iters, uint32_t a, double da, double db) {
    if ((threadIdx.x >> 5) & 1) {
        double x[8];
        for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k;
        for (int i = 0; i < iters; i++) dfma_body(x, da, db);
        double s = 0;
        for (int k = 0; k < 8; k++) s += x[k];
        out[blockIdx.x * blockDim.x + threadIdx.x] = (uint64_t)s;
    } else {
        uint64_t x[8];
        for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k;
        for (int i = 0; i < iters; i++) wide_body(x, a);
        uint64_t s = 0;
        for (int k = 0; k < 8; k++) s ^= x[k];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    }
}
template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best; }
    return best;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int blocks = p.multiProcessorCount * 8, threads = 256, iters = 2048;
    void* buf; cudaMalloc(&buf, (size_t)blocks * threads * 8);
    double ops = (double)blocks * threads * iters * 64.0;
    float t1 = timeit([&] { imad_k<<<blocks, threads>>>((uint32_t*)buf, iters, 0x9e3779b1u, 12345u); });
    float t2 = timeit([&] { wide_k<<<blocks, threads>>>((uint64_t*)buf, iters, 0x9e3779b1u); });
    float t3 = timeit([&] { dfma_k<<<blocks, threads>>>((double*)buf, iters, 1.0000001, 0.5); });
    float t4 = timeit([&] { mixed_k<<<blocks, threads>>>((uint64_t*)buf, iters, 0x9e3779b1u, 1.0000001, 0.5); });
    printf("{\"sms\": %d, \"imad_T\": %.2f, \"imad_wide_T\": %.2f, \"dfma_T\": %.2f, \"mixed_ms\": %.3f, \"wide_ms\": %.3f, \"dfma_ms\": %.3f, "
           "\"mixed_note\": \"half the warps each; if pipes are independent mixed_ms ~ max(wide_ms, dfma_ms)/2\"}\n",
           p.multiProcessorCount, ops / t1 / 1e9, ops / t2 / 1e9, ops / t3 / 1e9, t4, t2, t3);
    return 0;
}
