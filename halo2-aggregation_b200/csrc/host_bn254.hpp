// host_bn254.hpp — host-side BN254 arithmetic used by the product's glue code only:
// the last step of an MSM (Horner over a handful of window sums + one inversion), the
// transcript scalars of the verifier glue, and sums of per-GPU partial results.
// Everything data-parallel runs in the CUDA kernels; nothing here is a CPU fallback for them.
//
// Layout contract (include/h2agg.h): field elements are 4 x u64 little-endian limbs in
// Montgomery form (R = 2^256) — the in-memory form of the dependency types the reference uses
// (`bn256::Fr`, `G1Affine`; examples/simple-example.rs:552-553).
#pragma once
#include <cstdint>
#include <cstring>

namespace h2a_host {

typedef unsigned __int128 u128;

struct Mod {
    uint64_t p[4];
    uint64_t one[4];  // R mod p
    uint64_t r2[4];   // R^2 mod p
    uint64_t inv;     // -p^-1 mod 2^64
};

static const Mod MOD_Q = {{0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
                          {0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full},
                          {0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full},
                          0x87d20782e4866389ull};
static const Mod MOD_R = {{0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
                          {0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full},
                          {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull},
                          0xc2e1f593efffffffull};

// A value of either field; the modulus is passed explicitly (two instantiations below).
struct El {
    uint64_t v[4];
};

inline bool el_is_zero(const El& a) { return (a.v[0] | a.v[1] | a.v[2] | a.v[3]) == 0; }
inline bool el_eq(const El& a, const El& b) { return memcmp(a.v, b.v, 32) == 0; }

inline bool ge_p(const uint64_t* a, const Mod& m) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] != m.p[i]) return a[i] > m.p[i];
    }
    return true;
}
inline void sub_p(uint64_t* a, const Mod& m) {
    uint64_t borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a[i] - m.p[i] - borrow;
        a[i] = (uint64_t)d;
        borrow = (uint64_t)(d >> 127);
    }
}
inline El el_add(const El& a, const El& b, const Mod& m) {
    El r;
    uint64_t carry = 0;
    for (int i = 0; i < 4; i++) {
        u128 s = (u128)a.v[i] + b.v[i] + carry;
        r.v[i] = (uint64_t)s;
        carry = (uint64_t)(s >> 64);
    }
    if (carry || ge_p(r.v, m)) sub_p(r.v, m);
    return r;
}
inline El el_sub(const El& a, const El& b, const Mod& m) {
    El r;
    uint64_t borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a.v[i] - b.v[i] - borrow;
        r.v[i] = (uint64_t)d;
        borrow = (uint64_t)(d >> 127);
    }
    if (borrow) {
        uint64_t carry = 0;
        for (int i = 0; i < 4; i++) {
            u128 s = (u128)r.v[i] + m.p[i] + carry;
            r.v[i] = (uint64_t)s;
            carry = (uint64_t)(s >> 64);
        }
    }
    return r;
}
// Montgomery product by separate multiply (8 limbs) then 4 reduction rounds (SOS form).
inline El el_mul(const El& a, const El& b, const Mod& m) {
    uint64_t t[9] = {0};
    for (int i = 0; i < 4; i++) {
        uint64_t carry = 0;
        for (int j = 0; j < 4; j++) {
            u128 s = (u128)a.v[i] * b.v[j] + t[i + j] + carry;
            t[i + j] = (uint64_t)s;
            carry = (uint64_t)(s >> 64);
        }
        t[i + 4] = carry;
    }
    for (int i = 0; i < 4; i++) {
        uint64_t k = t[i] * m.inv, carry = 0;
        for (int j = 0; j < 4; j++) {
            u128 s = (u128)k * m.p[j] + t[i + j] + carry;
            t[i + j] = (uint64_t)s;
            carry = (uint64_t)(s >> 64);
        }
        for (int j = i + 4; carry && j < 9; j++) {
            u128 s = (u128)t[j] + carry;
            t[j] = (uint64_t)s;
            carry = (uint64_t)(s >> 64);
        }
    }
    El r = {{t[4], t[5], t[6], t[7]}};
    if (t[8] || ge_p(r.v, m)) sub_p(r.v, m);
    return r;
}
inline El el_one(const Mod& m) { El r; memcpy(r.v, m.one, 32); return r; }
inline El el_zero() { El r = {{0, 0, 0, 0}}; return r; }
inline El el_neg(const El& a, const Mod& m) { return el_is_zero(a) ? a : el_sub(el_zero(), a, m); }
inline El el_from_raw(const uint64_t raw[4], const Mod& m) {  // canonical -> Montgomery
    El a, r2;
    memcpy(a.v, raw, 32); memcpy(r2.v, m.r2, 32);
    return el_mul(a, r2, m);
}
inline El el_from_u64(uint64_t x, const Mod& m) { uint64_t raw[4] = {x, 0, 0, 0}; return el_from_raw(raw, m); }
inline void el_to_raw(const El& a, uint64_t raw[4], const Mod& m) {  // Montgomery -> canonical
    El o = {{1, 0, 0, 0}};
    El r = el_mul(a, o, m);
    memcpy(raw, r.v, 32);
}
inline El el_pow(const El& a, const uint64_t e[4], const Mod& m) {
    El acc = el_one(m);
    bool started = false;
    for (int i = 255; i >= 0; i--) {
        if (started) acc = el_mul(acc, acc, m);
        if ((e[i >> 6] >> (i & 63)) & 1) { acc = started ? el_mul(acc, a, m) : a; started = true; }
    }
    return acc;
}
inline El el_pow_u64(const El& a, uint64_t e, const Mod& m) { uint64_t t[4] = {e, 0, 0, 0}; return el_pow(a, t, m); }
inline El el_inv(const El& a, const Mod& m) {  // 0 -> 0
    if (el_is_zero(a)) return a;
    uint64_t e[4] = {m.p[0] - 2, m.p[1], m.p[2], m.p[3]};
    return el_pow(a, e, m);
}

// Typed wrappers
struct Fq { El e; };
struct Fr { El e; };
#define H2A_FIELD_OPS(T, M)                                                                    \
    inline T operator+(const T& a, const T& b) { return T{el_add(a.e, b.e, M)}; }               \
    inline T operator-(const T& a, const T& b) { return T{el_sub(a.e, b.e, M)}; }               \
    inline T operator*(const T& a, const T& b) { return T{el_mul(a.e, b.e, M)}; }               \
    inline bool operator==(const T& a, const T& b) { return el_eq(a.e, b.e); }                  \
    inline T neg(const T& a) { return T{el_neg(a.e, M)}; }                                      \
    inline T sqr(const T& a) { return T{el_mul(a.e, a.e, M)}; }                                 \
    inline T dbl(const T& a) { return T{el_add(a.e, a.e, M)}; }                                 \
    inline T inv(const T& a) { return T{el_inv(a.e, M)}; }                                      \
    inline bool is_zero(const T& a) { return el_is_zero(a.e); }                                 \
    inline T pow_u64(const T& a, uint64_t e) { return T{el_pow_u64(a.e, e, M)}; }
H2A_FIELD_OPS(Fq, MOD_Q)
H2A_FIELD_OPS(Fr, MOD_R)
#undef H2A_FIELD_OPS
inline Fq fq_one() { return Fq{el_one(MOD_Q)}; }
inline Fq fq_zero() { return Fq{el_zero()}; }
inline Fr fr_one() { return Fr{el_one(MOD_R)}; }
inline Fr fr_zero() { return Fr{el_zero()}; }
inline Fr fr_from_u64(uint64_t x) { return Fr{el_from_u64(x, MOD_R)}; }
inline Fr fr_load(const uint8_t* p) { Fr r; memcpy(r.e.v, p, 32); return r; }
inline void fr_store(uint8_t* p, const Fr& a) { memcpy(p, a.e.v, 32); }
inline void fr_to_raw(const Fr& a, uint64_t raw[4]) { el_to_raw(a.e, raw, MOD_R); }
inline void fq_to_raw(const Fq& a, uint64_t raw[4]) { el_to_raw(a.e, raw, MOD_Q); }
inline Fr fr_from_raw(const uint64_t raw[4]) { return Fr{el_from_raw(raw, MOD_R)}; }
inline Fq fq_from_raw(const uint64_t raw[4]) { return Fq{el_from_raw(raw, MOD_Q)}; }

// omega_k = ROOT_OF_UNITY^(2^(28-k)); ROOT_OF_UNITY = 7^((r-1)/2^28)
inline Fr fr_root_of_unity(int k) {
    static const uint64_t ROOT_RAW[4] = {0xd34f1ed960c37c9cull, 0x3215cf6dd39329c8ull, 0x98865ea93dd31f74ull,
                                         0x03ddb9f5166d18b7ull};
    Fr w = fr_from_raw(ROOT_RAW);
    for (int i = k; i < 28; i++) w = sqr(w);
    return w;
}

// ------------------------------------------------------------------ G1, XYZZ coordinates on the host
struct PointA {  // affine, identity = (0,0)
    Fq x, y;
};
inline bool is_identity(const PointA& a) { return is_zero(a.x) && is_zero(a.y); }
inline PointA affine_load(const uint8_t* p) { PointA a; memcpy(a.x.e.v, p, 32); memcpy(a.y.e.v, p + 32, 32); return a; }
inline void affine_store(uint8_t* p, const PointA& a) { memcpy(p, a.x.e.v, 32); memcpy(p + 32, a.y.e.v, 32); }

struct PointX {  // x = X/ZZ, y = Y/ZZZ; identity <=> ZZ == 0
    Fq x, y, zz, zzz;
};
inline PointX px_identity() { return PointX{fq_zero(), fq_zero(), fq_zero(), fq_zero()}; }
inline bool is_identity(const PointX& a) { return is_zero(a.zz); }
inline PointX px_from_affine(const PointA& a) {
    if (is_identity(a)) return px_identity();
    return PointX{a.x, a.y, fq_one(), fq_one()};
}
inline PointX px_load(const uint8_t* p) {
    PointX r;
    memcpy(r.x.e.v, p, 32); memcpy(r.y.e.v, p + 32, 32); memcpy(r.zz.e.v, p + 64, 32); memcpy(r.zzz.e.v, p + 96, 32);
    return r;
}
inline PointX px_dbl(const PointX& a) {
    if (is_identity(a)) return a;
    Fq u = dbl(a.y), v = sqr(u), w = u * v, s = a.x * v, xx = sqr(a.x), m = dbl(xx) + xx;
    PointX r;
    r.x = sqr(m) - dbl(s);
    r.y = m * (s - r.x) - w * a.y;
    r.zz = v * a.zz;
    r.zzz = w * a.zzz;
    return r;
}
inline PointX px_add(const PointX& a, const PointX& b) {
    if (is_identity(a)) return b;
    if (is_identity(b)) return a;
    Fq u1 = a.x * b.zz, u2 = b.x * a.zz, s1 = a.y * b.zzz, s2 = b.y * a.zzz;
    Fq p = u2 - u1, r = s2 - s1;
    if (is_zero(p)) return is_zero(r) ? px_dbl(a) : px_identity();
    Fq pp = sqr(p), ppp = p * pp, q = u1 * pp;
    PointX o;
    o.x = sqr(r) - ppp - dbl(q);
    o.y = r * (q - o.x) - s1 * ppp;
    o.zz = a.zz * b.zz * pp;
    o.zzz = a.zzz * b.zzz * ppp;
    return o;
}
inline PointX px_neg(const PointX& a) { return PointX{a.x, neg(a.y), a.zz, a.zzz}; }
inline PointA px_to_affine(const PointX& a) {
    if (is_identity(a)) return PointA{fq_zero(), fq_zero()};
    Fq zi = inv(a.zzz);               // 1/ZZZ
    Fq zzi = sqr(zi * a.zz);          // (ZZ/ZZZ)^2 = 1/ZZ   (since ZZ^3 = ZZZ^2)
    return PointA{a.x * zzi, a.y * zi};
}
inline PointX px_mul(const PointX& a, const Fr& s) {  // double-and-add over the canonical scalar
    uint64_t e[4];
    fr_to_raw(s, e);
    PointX acc = px_identity();
    for (int i = 255; i >= 0; i--) {
        acc = px_dbl(acc);
        if ((e[i >> 6] >> (i & 63)) & 1) acc = px_add(acc, a);
    }
    return acc;
}
inline bool on_curve(const PointA& a) {
    if (is_identity(a)) return true;
    return sqr(a.y) == sqr(a.x) * a.x + Fq{el_from_u64(3, MOD_Q)};
}

}  // namespace h2a_host
