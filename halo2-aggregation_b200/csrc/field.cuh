// field.cuh — BN254 Fq / Fr arithmetic for sm_100a: 8 x 32-bit limbs, Montgomery form (R = 2^256).
//
// The in-memory form is the one a Rust `&[Fr]` / `G1Affine` already has (4 x u64 LE Montgomery
// limbs == 8 x u32 LE), so buffers cross the C ABI untouched (include/h2agg.h).
// Replaces, for the hot path, the dependency arithmetic the reference reaches through
// `halo2::pairing::bn256::{Fr, G1Affine}` (examples/simple-example.rs:552-553).
//
// Multiplication is a row-interleaved (CIOS-style) Montgomery product in which the partial
// products of even and odd limbs of `a` accumulate in two separate 8-limb accumulators, so every
// `mad.lo.cc` / `madc.hi.cc` pair lands on an aligned register pair and ptxas emits one
// IMAD.WIDE.U32(.X) per 32x32 product with the carry riding the predicate chain.  Every carry
// chain is a single asm statement: the CC flag never lives across statement boundaries.
#pragma once
#include <cstdint>
#ifdef __CUDACC__
#include "field_mul_gen.cuh"
#endif
// H2A_WIDE_SQR (default 1): sqr() is the generated dedicated square of field_mul_gen.cuh (100 IMAD.WIDE instead of 136:
// 76.6 against 65.0 G/s in tools/micro/mulvar.cu); 0 squares with the product.  The generated Karatsuba product (112 IMAD.WIDE
// but 80 more plain instructions) measured slower than the row-interleaved product below and is not used.
#ifndef H2A_WIDE_SQR
#define H2A_WIDE_SQR 1
#endif

namespace h2a {

enum FieldId { FQ = 0, FR = 1 };

template <int F> struct FieldConst;

template <> struct FieldConst<FQ> {
    __host__ __device__ static constexpr uint32_t mod(int i) {
        constexpr uint32_t m[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u,
                                   0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return m[i];
    }
    __host__ __device__ static constexpr uint32_t one(int i) {  // R mod p
        constexpr uint32_t m[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u,
                                   0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return m[i];
    }
    __host__ __device__ static constexpr uint32_t r2(int i) {  // R^2 mod p
        constexpr uint32_t m[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u,
                                   0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
        return m[i];
    }
    static constexpr uint32_t INV = 0xe4866389u;  // -p^-1 mod 2^32
};
template <> struct FieldConst<FR> {
    __host__ __device__ static constexpr uint32_t mod(int i) {
        constexpr uint32_t m[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                                   0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return m[i];
    }
    __host__ __device__ static constexpr uint32_t one(int i) {
        constexpr uint32_t m[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u,
                                   0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return m[i];
    }
    __host__ __device__ static constexpr uint32_t r2(int i) {
        constexpr uint32_t m[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u,
                                   0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
        return m[i];
    }
    static constexpr uint32_t INV = 0xefffffffu;
};

#ifdef __CUDACC__

// acc[0..7] += (x0,x2,x4,x6 as 64-bit products with b, at limb offsets 0,2,4,6); top += carry out
__device__ __forceinline__ void mad_even_row(uint32_t* acc, uint32_t& top, uint32_t x0, uint32_t x2,
                                             uint32_t x4, uint32_t x6, uint32_t b) {
    asm("mad.lo.cc.u32  %0, %9,  %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9,  %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32       %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
          "+r"(acc[6]), "+r"(acc[7]), "+r"(top)
        : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(b));
}

// same, without a carry out (caller guarantees none)
__device__ __forceinline__ void mad_row_nocarry(uint32_t* acc, uint32_t x0, uint32_t x2, uint32_t x4,
                                                uint32_t x6, uint32_t b) {
    asm("mad.lo.cc.u32  %0, %8,  %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8,  %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9,  %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9,  %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32    %7, %11, %12, %7;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
          "+r"(acc[6]), "+r"(acc[7])
        : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(b));
}

// x0 += y[1] (carry c); then y <- (y >> 64 bits) + (x1,x3,x5,x7 products with b) + c, i.e. the
// accumulator that was aligned at limb 0 (and whose limb 0 is now zero) is re-based to limb 1.
__device__ __forceinline__ void shift_mad_odd_row(uint32_t& x0, uint32_t* y, uint32_t a1, uint32_t a3,
                                                  uint32_t a5, uint32_t a7, uint32_t b) {
    asm("add.cc.u32     %0, %0, %2;\n\t"
        "madc.lo.cc.u32 %1, %9,  %13, %3;\n\t"
        "madc.hi.cc.u32 %2, %9,  %13, %4;\n\t"
        "madc.lo.cc.u32 %3, %10, %13, %5;\n\t"
        "madc.hi.cc.u32 %4, %10, %13, %6;\n\t"
        "madc.lo.cc.u32 %5, %11, %13, %7;\n\t"
        "madc.hi.cc.u32 %6, %11, %13, %8;\n\t"
        "madc.lo.cc.u32 %7, %12, %13, 0;\n\t"
        "madc.hi.u32    %8, %12, %13, 0;"
        : "+r"(x0), "+r"(y[0]), "+r"(y[1]), "+r"(y[2]), "+r"(y[3]), "+r"(y[4]), "+r"(y[5]), "+r"(y[6]),
          "+r"(y[7])
        : "r"(a1), "r"(a3), "r"(a5), "r"(a7), "r"(b));
}

template <int F>
struct Fp {
    uint32_t l[8];
    typedef FieldConst<F> C;

    __device__ __forceinline__ static Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = 0;
        return r;
    }
    __device__ __forceinline__ static Fp one() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = C::one(i);
        return r;
    }
    __device__ __forceinline__ static Fp r2() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = C::r2(i);
        return r;
    }
    __device__ __forceinline__ static Fp load(const void* p) {  // 32-byte aligned
        Fp r;
        uint4 a = ((const uint4*)p)[0], b = ((const uint4*)p)[1];
        r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
        r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
        return r;
    }
    // gather load for random table lookups: one 256-bit LDG (sm_100) instead of two 128-bit ones.  Measured on B200
    // (tools/micro/gather.cu, 3.5 GB table): 38.0 G lookups/s of 32 bytes against 32.3 with 2 x LDG.128, 23.8 against
    // 17.7 for 64-byte lookups; L2-only (.cg) 128-bit loads are slower than either (19.1 / 9.6).
    __device__ __forceinline__ static Fp load_gather(const void* p) {  // 32-byte aligned, read-only data
        Fp r;
        asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7])
                     : "l"(p));
        return r;
    }
    __device__ __forceinline__ void store(void* p) const {
        ((uint4*)p)[0] = make_uint4(l[0], l[1], l[2], l[3]);
        ((uint4*)p)[1] = make_uint4(l[4], l[5], l[6], l[7]);
    }
    __device__ __forceinline__ bool is_zero() const {
        return (l[0] | l[1] | l[2] | l[3] | l[4] | l[5] | l[6] | l[7]) == 0;
    }
    __device__ __forceinline__ bool operator==(const Fp& o) const {
        uint32_t d = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) d |= l[i] ^ o.l[i];
        return d == 0;
    }

    // r = v - p if v >= p else v  (v < 2p)
    __device__ __forceinline__ static void reduce_once(uint32_t* v) {
        uint32_t t[8], borrow;
        asm("sub.cc.u32  %0, %9,  %17;\n\t"
            "subc.cc.u32 %1, %10, %18;\n\t"
            "subc.cc.u32 %2, %11, %19;\n\t"
            "subc.cc.u32 %3, %12, %20;\n\t"
            "subc.cc.u32 %4, %13, %21;\n\t"
            "subc.cc.u32 %5, %14, %22;\n\t"
            "subc.cc.u32 %6, %15, %23;\n\t"
            "subc.cc.u32 %7, %16, %24;\n\t"
            "subc.u32    %8, 0, 0;"
            : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]),
              "=r"(borrow)
            : "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
              "r"(C::mod(0)), "r"(C::mod(1)), "r"(C::mod(2)), "r"(C::mod(3)), "r"(C::mod(4)), "r"(C::mod(5)),
              "r"(C::mod(6)), "r"(C::mod(7)));
        if (borrow == 0) {
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = t[i];
        }
    }

    __device__ __forceinline__ Fp operator+(const Fp& o) const {
        Fp r;
        asm("add.cc.u32  %0, %8,  %16;\n\t"
            "addc.cc.u32 %1, %9,  %17;\n\t"
            "addc.cc.u32 %2, %10, %18;\n\t"
            "addc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\t"
            "addc.cc.u32 %5, %13, %21;\n\t"
            "addc.cc.u32 %6, %14, %22;\n\t"
            "addc.u32    %7, %15, %23;"
            : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]),
              "=r"(r.l[7])
            : "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]), "r"(l[4]), "r"(l[5]), "r"(l[6]), "r"(l[7]), "r"(o.l[0]),
              "r"(o.l[1]), "r"(o.l[2]), "r"(o.l[3]), "r"(o.l[4]), "r"(o.l[5]), "r"(o.l[6]), "r"(o.l[7]));
        reduce_once(r.l);  // a,b < p < 2^254: the sum fits 255 bits
        return r;
    }
    __device__ __forceinline__ Fp operator-(const Fp& o) const {
        Fp r;
        uint32_t borrow;
        asm("sub.cc.u32  %0, %9,  %17;\n\t"
            "subc.cc.u32 %1, %10, %18;\n\t"
            "subc.cc.u32 %2, %11, %19;\n\t"
            "subc.cc.u32 %3, %12, %20;\n\t"
            "subc.cc.u32 %4, %13, %21;\n\t"
            "subc.cc.u32 %5, %14, %22;\n\t"
            "subc.cc.u32 %6, %15, %23;\n\t"
            "subc.cc.u32 %7, %16, %24;\n\t"
            "subc.u32    %8, 0, 0;"
            : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]),
              "=r"(r.l[7]), "=r"(borrow)
            : "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]), "r"(l[4]), "r"(l[5]), "r"(l[6]), "r"(l[7]), "r"(o.l[0]),
              "r"(o.l[1]), "r"(o.l[2]), "r"(o.l[3]), "r"(o.l[4]), "r"(o.l[5]), "r"(o.l[6]), "r"(o.l[7]));
        // borrow is 0 or 0xffffffff: add back (p & borrow)
        asm("add.cc.u32  %0, %0, %8;\n\t"
            "addc.cc.u32 %1, %1, %9;\n\t"
            "addc.cc.u32 %2, %2, %10;\n\t"
            "addc.cc.u32 %3, %3, %11;\n\t"
            "addc.cc.u32 %4, %4, %12;\n\t"
            "addc.cc.u32 %5, %5, %13;\n\t"
            "addc.cc.u32 %6, %6, %14;\n\t"
            "addc.u32    %7, %7, %15;"
            : "+r"(r.l[0]), "+r"(r.l[1]), "+r"(r.l[2]), "+r"(r.l[3]), "+r"(r.l[4]), "+r"(r.l[5]), "+r"(r.l[6]),
              "+r"(r.l[7])
            : "r"(C::mod(0) & borrow), "r"(C::mod(1) & borrow), "r"(C::mod(2) & borrow), "r"(C::mod(3) & borrow),
              "r"(C::mod(4) & borrow), "r"(C::mod(5) & borrow), "r"(C::mod(6) & borrow), "r"(C::mod(7) & borrow));
        return r;
    }
    __device__ __forceinline__ Fp neg() const { return zero() - *this; }
    __device__ __forceinline__ Fp dbl() const { return *this + *this; }

    // Montgomery product a*b/R mod p, fully reduced.  Inputs < p.
    __device__ __forceinline__ Fp operator*(const Fp& o) const {
        return mul_cios(o);
    }
    __device__ __forceinline__ Fp sqr() const {
#if H2A_WIDE_SQR
        Fp r;
        mont_sqr_wide<F>(r.l, l);
        reduce_once(r.l);
        return r;
#else
        return mul_cios(*this);
#endif
    }
    // the row-interleaved (CIOS) product: 136 IMAD.WIDE, fewest other instructions
    __device__ __forceinline__ Fp mul_cios(const Fp& o) const {
        const uint32_t* a = l;
        uint32_t X[8], Y[8];  // X aligned at limb 0, Y aligned at limb 1 (roles swap every row)
        {   // row 0
            uint32_t b = o.l[0];
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
                uint64_t pe = (uint64_t)a[j] * b, po = (uint64_t)a[j + 1] * b;
                X[j] = (uint32_t)pe; X[j + 1] = (uint32_t)(pe >> 32);
                Y[j] = (uint32_t)po; Y[j + 1] = (uint32_t)(po >> 32);
            }
            uint32_t m = X[0] * C::INV;
            mad_row_nocarry(Y, C::mod(1), C::mod(3), C::mod(5), C::mod(7), m);
            mad_even_row(X, Y[7], C::mod(0), C::mod(2), C::mod(4), C::mod(6), m);
        }
#pragma unroll
        for (int i = 1; i < 8; i++) {
            uint32_t b = o.l[i];
            // after the previous row X[0] == 0; old Y becomes the limb-0 accumulator
            uint32_t* nx = (i & 1) ? Y : X;  // new limb-0 accumulator
            uint32_t* ny = (i & 1) ? X : Y;  // re-based to limb 1
            shift_mad_odd_row(nx[0], ny, a[1], a[3], a[5], a[7], b);
            mad_even_row(nx, ny[7], a[0], a[2], a[4], a[6], b);
            uint32_t m = nx[0] * C::INV;
            mad_row_nocarry(ny, C::mod(1), C::mod(3), C::mod(5), C::mod(7), m);
            mad_even_row(nx, ny[7], C::mod(0), C::mod(2), C::mod(4), C::mod(6), m);
        }
        // after 8 rows the limb-0 accumulator is Y (i=7 odd -> nx = Y) with Y[0]==0, limb-1 one is X
        Fp r;
        asm("add.cc.u32  %0, %8,  %16;\n\t"
            "addc.cc.u32 %1, %9,  %17;\n\t"
            "addc.cc.u32 %2, %10, %18;\n\t"
            "addc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\t"
            "addc.cc.u32 %5, %13, %21;\n\t"
            "addc.cc.u32 %6, %14, %22;\n\t"
            "addc.u32    %7, %15, 0;"
            : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]),
              "=r"(r.l[7])
            : "r"(X[0]), "r"(X[1]), "r"(X[2]), "r"(X[3]), "r"(X[4]), "r"(X[5]), "r"(X[6]), "r"(X[7]), "r"(Y[1]),
              "r"(Y[2]), "r"(Y[3]), "r"(Y[4]), "r"(Y[5]), "r"(Y[6]), "r"(Y[7]));
        reduce_once(r.l);
        return r;
    }
    // Montgomery -> canonical integer limbs (multiply by raw 1)
    __device__ __forceinline__ Fp from_mont() const {
        Fp o = zero();
        o.l[0] = 1;
        return (*this) * o;
    }
    __device__ __forceinline__ Fp to_mont() const { return (*this) * r2(); }  // canonical -> Montgomery

    __device__ Fp pow_limbs(const uint32_t* e, int nbits) const {  // e little-endian limbs
        Fp acc = one();
        for (int i = nbits - 1; i >= 0; i--) {
            acc = acc.sqr();
            if ((e[i >> 5] >> (i & 31)) & 1) acc = acc * (*this);
        }
        return acc;
    }
    __device__ Fp inv_fermat() const {  // a^(p-2), 0 -> 0; kept as the cross-check of inv()
        uint32_t e[8];
#pragma unroll
        for (int i = 0; i < 8; i++) e[i] = C::mod(i);
        e[0] -= 2;
        return pow_limbs(e, 254);
    }
    // ---- inversion by "safegcd" division steps (Bernstein-Yang, half-delta variant): 20 batches of 30 branch-free
    // steps on the low words of (f, g) = (p, a), each batch followed by one 2x2-matrix update of the full-width
    // (f, g) and of the Bezout pair (d, e) mod p.  Values live in nine signed 30-bit limbs so that every partial
    // sum fits a 64-bit accumulator.  600 steps always reach g = 0, f = +-1 for a 256-bit modulus; no data-dependent
    // branch, so a warp never diverges, and ~13 k plain integer instructions replace the ~350 dependent Montgomery
    // products of Fermat.  Input and output in Montgomery form: inv(aR) = a^-1 R^-1 as integers, times R^3 / R
    // = a^-1 R.  0 -> 0.
    static constexpr int32_t M30 = (1 << 30) - 1;
    __host__ __device__ static constexpr int32_t mod30(int i) {  // limb i of the modulus in base 2^30
        return (int32_t)((((uint64_t)C::mod((30 * i) >> 5) | ((uint64_t)(((30 * i) >> 5) + 1 < 8 ? C::mod(((30 * i) >> 5) + 1) : 0u) << 32)) >>
                          ((30 * i) & 31)) & (uint64_t)M30);
    }
    __host__ __device__ static constexpr uint32_t modinv30() {   // p^-1 mod 2^30 (Newton on the low word)
        uint32_t m = C::mod(0), x = m;
        for (int k = 0; k < 5; k++) x *= 2u - m * x;
        return x & (uint32_t)M30;
    }
    __device__ __noinline__ Fp inv() const {
        if (is_zero()) return *this;
        int32_t d[9], e[9], f[9], g[9];
#pragma unroll
        for (int i = 0; i < 9; i++) {
            const int w = (30 * i) >> 5, sh = (30 * i) & 31;
            const uint64_t two = (uint64_t)l[w] | ((uint64_t)(w + 1 < 8 ? l[w + 1] : 0u) << 32);
            g[i] = (int32_t)((two >> sh) & (uint64_t)M30);
            f[i] = mod30(i);
            d[i] = 0;
            e[i] = 0;
        }
        e[0] = 1;
        int32_t zeta = -1;   // -(delta + 1/2)
#pragma unroll 1
        for (int batch = 0; batch < 20; batch++) {
            // 30 division steps on the low limbs; (u v; q r) / 2^30 is the transition matrix of the batch
            int32_t u = 1, v = 0, q = 0, r = 1;
            uint32_t fl = (uint32_t)f[0], gl = (uint32_t)g[0];
#pragma unroll 6
            for (int step = 0; step < 30; step++) {
                int32_t c1 = zeta >> 31;
                const int32_t c2 = -(int32_t)(gl & 1u);
                const uint32_t x = (fl ^ (uint32_t)c1) - (uint32_t)c1;
                const int32_t y = (u ^ c1) - c1, z = (v ^ c1) - c1;
                gl += x & (uint32_t)c2;
                q += y & c2;
                r += z & c2;
                c1 &= c2;
                zeta = (zeta ^ c1) - 1;
                fl += gl & (uint32_t)c1;
                u += q & c1;
                v += r & c1;
                gl >>= 1;
                u <<= 1;
                v <<= 1;
            }
            {   // (d, e) <- (u d + v e, q d + r e) / 2^30 mod p, kept in (-2p, p)
                const int32_t sd = d[8] >> 31, se = e[8] >> 31;
                int32_t md = (u & sd) + (v & se), me = (q & sd) + (r & se);
                long long cd = (long long)u * d[0] + (long long)v * e[0];
                long long ce = (long long)q * d[0] + (long long)r * e[0];
                md -= (int32_t)((modinv30() * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
                me -= (int32_t)((modinv30() * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
                cd += (long long)mod30(0) * md;
                ce += (long long)mod30(0) * me;
                cd >>= 30;
                ce >>= 30;
#pragma unroll
                for (int i = 1; i < 9; i++) {
                    cd += (long long)u * d[i] + (long long)v * e[i] + (long long)mod30(i) * md;
                    ce += (long long)q * d[i] + (long long)r * e[i] + (long long)mod30(i) * me;
                    d[i - 1] = (int32_t)cd & M30;
                    e[i - 1] = (int32_t)ce & M30;
                    cd >>= 30;
                    ce >>= 30;
                }
                d[8] = (int32_t)cd;
                e[8] = (int32_t)ce;
            }
            {   // (f, g) <- (u f + v g, q f + r g) / 2^30 (exact)
                long long cf = (long long)u * f[0] + (long long)v * g[0];
                long long cg = (long long)q * f[0] + (long long)r * g[0];
                cf >>= 30;
                cg >>= 30;
#pragma unroll
                for (int i = 1; i < 9; i++) {
                    cf += (long long)u * f[i] + (long long)v * g[i];
                    cg += (long long)q * f[i] + (long long)r * g[i];
                    f[i - 1] = (int32_t)cf & M30;
                    g[i - 1] = (int32_t)cg & M30;
                    cf >>= 30;
                    cg >>= 30;
                }
                f[8] = (int32_t)cf;
                g[8] = (int32_t)cg;
            }
        }
        // now g = 0 and f = +-1: a^-1 = sign(f) * d, brought into [0, p)
        {
            int32_t add = d[8] >> 31;
            const int32_t neg = f[8] >> 31;
#pragma unroll
            for (int i = 0; i < 9; i++) {
                d[i] += mod30(i) & add;
                d[i] = (d[i] ^ neg) - neg;
            }
#pragma unroll
            for (int i = 0; i < 8; i++) {
                d[i + 1] += d[i] >> 30;
                d[i] &= M30;
            }
            add = d[8] >> 31;
#pragma unroll
            for (int i = 0; i < 9; i++) d[i] += mod30(i) & add;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                d[i + 1] += d[i] >> 30;
                d[i] &= M30;
            }
        }
        Fp out;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            const int i = (32 * w) / 30, o = 32 * w - 30 * i;
            uint32_t word = ((uint32_t)d[i] >> o) | ((uint32_t)d[i + 1] << (30 - o));
            out.l[w] = word;
        }
        return out * (r2() * r2());
    }

    // is canonical-integer limb vector v >= modulus ?
    __device__ __forceinline__ static bool geq_mod(const uint32_t* v) {
#pragma unroll
        for (int i = 7; i >= 0; i--) {
            if (v[i] > C::mod(i)) return true;
            if (v[i] < C::mod(i)) return false;
        }
        return true;
    }
};

typedef Fp<FQ> Fq;
typedef Fp<FR> Fr;

#endif  // __CUDACC__
}  // namespace h2a
