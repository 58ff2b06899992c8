// Random-gather microbenchmark: 32-byte and 64-byte lookups into a table much larger than L2, with the load
// flavours a table-driven MSM could use.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather gather.cu
// Prints the rate of useful bytes for each flavour (B200: see tools/micro/README.md).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
template <int MODE>
__device__ __forceinline__ uint32_t load32(const uint8_t* p) {   // 32 bytes -> xor of the words
    uint32_t r[8];
    if (MODE == 0) {          // 2 x LDG.128, read-only path through L1 (what `const __restrict__` gives)
        uint4 a = __ldg((const uint4*)p), b = __ldg((const uint4*)p + 1);
        return a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w;
    } else if (MODE == 1) {   // 2 x LDG.128, L2 only
        uint4 a = __ldcg((const uint4*)p), b = __ldcg((const uint4*)p + 1);
        return a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w;
    } else if (MODE == 2) {   // LDG.256 read-only path
        asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]) : "l"(p));
    } else if (MODE == 3) {   // LDG.256, L2 only
        asm volatile("ld.global.cg.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]) : "l"(p));
    } else {                  // LDG.256 read-only path, no L1 allocation
        asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]) : "l"(p));
    }
    return r[0] ^ r[1] ^ r[2] ^ r[3] ^ r[4] ^ r[5] ^ r[6] ^ r[7];
}
template <int MODE, int BYTES>
__global__ void gather_kernel(const uint8_t* table, uint32_t entries, int per_thread, uint32_t* out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
#pragma unroll 4
    for (int i = 0; i < per_thread; i++) {
        const uint32_t e = mix(t * 977u + i * 0x9e3779b9u) % entries;
        const uint8_t* p = table + 64ull * e;
        acc ^= load32<MODE>(p);
        if (BYTES == 64) acc ^= load32<MODE>(p + 32);
    }
    out[t] = acc;
}
template <int MODE, int BYTES>
static void run(const char* name, const uint8_t* table, uint32_t entries, uint32_t* out) {
    const int threads = 148 * 2048, per = 64;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    gather_kernel<MODE, BYTES><<<threads / 128, 128>>>(table, entries, per, out);
    cudaEventRecord(a);
    for (int r = 0; r < 3; r++) gather_kernel<MODE, BYTES><<<threads / 128, 128>>>(table, entries, per, out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double lookups = 3.0 * threads * per;
    printf("{\"load\": \"%s\", \"bytes_per_lookup\": %d, \"glookups_per_s\": %.2f, \"useful_gb_per_s\": %.1f}\n", name, BYTES,
           lookups / (ms * 1e-3) / 1e9, lookups * BYTES / (ms * 1e-3) / 1e9);
}
int main() {
    const uint32_t entries = 54525952u;   // 13 windows x 2^22 points, 64 bytes each = 3.5 GB
    uint8_t* table; uint32_t* out;
    if (cudaMalloc(&table, 64ull * entries) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&out, 4ull * 148 * 2048);
    cudaMemset(table, 1, 64ull * entries);
    run<0, 32>("2xLDG.128 nc (L1)", table, entries, out);
    run<1, 32>("2xLDG.128 cg (L2 only)", table, entries, out);
    run<2, 32>("LDG.256 nc (L1)", table, entries, out);
    run<3, 32>("LDG.256 cg (L2 only)", table, entries, out);
    run<4, 32>("LDG.256 nc no_allocate", table, entries, out);
    run<0, 64>("2xLDG.128 nc (L1)", table, entries, out);
    run<1, 64>("2xLDG.128 cg (L2 only)", table, entries, out);
    run<2, 64>("LDG.256 nc (L1)", table, entries, out);
    run<3, 64>("LDG.256 cg (L2 only)", table, entries, out);
    run<4, 64>("LDG.256 nc no_allocate", table, entries, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
