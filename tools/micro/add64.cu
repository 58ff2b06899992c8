#include <cstdint>
__global__ void k(uint64_t* o, const uint64_t* a, const uint64_t* b, const double* d) {
    int i = threadIdx.x;
    uint64_t x = a[i], y = b[i];
    double p = d[i], q = d[i + 32];
    double hi = __fma_rz(p, q, 0x1p104);
    double lo = __fma_rz(p, q, 0x1p104 + 0x1p52 - hi);
    uint64_t s = x + y + (uint64_t)__double_as_longlong(hi) + (uint64_t)__double_as_longlong(lo);
    o[i] = s + (s >> 52) + (x & 0xfffffffffffffull);
}
