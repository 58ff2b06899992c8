// Montgomery product variants of csrc/field.cuh: correctness of the generated Karatsuba product / square against the
// row-interleaved product, and the rate of each in a dependent chain (4 independent chains per thread).
// Build (the Karatsuba product is not in the shipped header; generate one that has it):
//   python ../gen_field_mul.py --with-mul --only-mul --out=_bin/field_mul_kara.cuh
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I_bin -I../../halo2-aggregation_b200/csrc -o _bin/mulvar mulvar.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "field.cuh"
#include "field_mul_kara.cuh"
using namespace h2a;
#ifndef H2A_HAVE_WIDE_MUL
#error "build against a header generated with --with-mul (see above)"
#endif
template <int F> __device__ __forceinline__ Fp<F> mul_kara(const Fp<F>& a, const Fp<F>& b) {
    Fp<F> r;
    mont_mul_wide<F>(r.l, a.l, b.l);
    Fp<F>::reduce_once(r.l);
    return r;
}

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
template <int F>
__device__ Fp<F> rnd(uint32_t seed) {
    Fp<F> r;
    for (int i = 0; i < 8; i++) r.l[i] = mix(seed * 8u + i);
    const uint32_t sel = mix(seed ^ 0xabcdu);
    for (int i = 0; i < 8; i++) {     // a third of the limbs all ones / zero
        const uint32_t k = (sel >> (2 * i)) & 3u;
        if (k == 1) r.l[i] = 0xffffffffu;
        if (k == 2 && (sel >> 20 & 1)) r.l[i] = 0;
    }
    r.l[7] &= 0x1fffffffu;    // < p
    return r;
}
template <int F>
__global__ void check_kernel(uint32_t n, unsigned long long* bad) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    Fp<F> a = rnd<F>(2 * t), b = rnd<F>(2 * t + 1);
    if (t == 0) { a = Fp<F>::zero(); }
    if (t == 1) { for (int i = 0; i < 8; i++) a.l[i] = FieldConst<F>::mod(i); a.l[0] -= 1; b = a; }
    Fp<F> want = a.mul_cios(b), got = mul_kara<F>(a, b);
    Fp<F> want2 = a.mul_cios(a), got2 = a.sqr();
    if (!(want == got)) atomicAdd(bad, 1ull);
    if (!(want2 == got2)) atomicAdd(bad + 1, 1ull);
}
template <int F, int MODE>
__global__ void __launch_bounds__(256) rate_kernel(int iters, uint32_t* out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    Fp<F> x[4], y = rnd<F>(t + 77);
    for (int k = 0; k < 4; k++) x[k] = rnd<F>(4 * t + k);
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (MODE == 0) x[k] = x[k].mul_cios(y);
            else if (MODE == 1) x[k] = mul_kara<F>(x[k], y);
            else x[k] = x[k].sqr();
        }
    }
    uint32_t acc = 0;
    for (int k = 0; k < 4; k++) for (int i = 0; i < 8; i++) acc ^= x[k].l[i];
    out[t] = acc;
}
template <int F, int MODE>
static void rate(const char* name, uint32_t* out) {
    const int blocks = 148 * 8, iters = 2000;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    rate_kernel<F, MODE><<<blocks, 256>>>(iters, out);
    cudaEventRecord(a);
    rate_kernel<F, MODE><<<blocks, 256>>>(iters, out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    printf("{\"variant\": \"%s\", \"field\": %d, \"giga_products_per_s\": %.2f}\n", name, F, 4.0 * blocks * 256 * iters / (ms * 1e-3) / 1e9);
}
int main() {
    unsigned long long* bad; uint32_t* out;
    cudaMalloc(&bad, 32); cudaMemset(bad, 0, 32);
    cudaMalloc(&out, 4ull * 148 * 8 * 256);
    const uint32_t n = 1u << 22;
    check_kernel<0><<<n / 256, 256>>>(n, bad);
    check_kernel<1><<<n / 256, 256>>>(n, bad + 2);
    unsigned long long h[4];
    cudaMemcpy(h, bad, 32, cudaMemcpyDeviceToHost);
    printf("{\"checked\": %u, \"fq_mul_bad\": %llu, \"fq_sqr_bad\": %llu, \"fr_mul_bad\": %llu, \"fr_sqr_bad\": %llu}\n", n, h[0], h[1], h[2], h[3]);
    rate<0, 0>("cios product (136 IMAD.WIDE)", out);
    rate<0, 1>("karatsuba product (112)", out);
    rate<0, 2>("square (100)", out);
    rate<1, 0>("cios product (136 IMAD.WIDE)", out);
    rate<1, 1>("karatsuba product (112)", out);
    rate<1, 2>("square (100)", out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return (h[0] | h[1] | h[2] | h[3]) ? 2 : 0;
}
