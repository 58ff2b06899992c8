// oracle/bn254.hpp — CPU restatement of BN254 ("bn256") Fq / Fr / G1 arithmetic.
//
// TEST INFRASTRUCTURE ONLY.  The product (halo2-aggregation_b200/) never includes,
// links or calls anything under oracle/; only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py do, as the checker / reported baseline.
//
// PARITY UNPINNED: the arithmetic the reference relies on lives in two un-vendored git
// branches (Cargo.toml:10 halo2wrong@agg2, Cargo.toml:12 halo2@kzg-agg2; no lockfile,
// .gitignore:7) and the reference has no tests or golden vectors (src/lib.rs:43-44).
// This restatement follows the published algorithms of zcash/halo2 v0.1.0-beta-era
// `arithmetic.rs` and `pairing::bn256` (4x64-bit Montgomery limbs, R = 2^256,
// Jacobian G1) and is pinned to the mathematical known answers of SURVEY.md App. A
// and to oracle/pymodel.py (Python big integers).  Reference call sites that consume
// these types: examples/simple-example.rs:552-553 (bn256::Fr, G1Affine), :638-640.
#pragma once
#include <cstdint>
#include <cstring>

namespace orc {

typedef unsigned __int128 u128;

struct FqParams {
    static constexpr uint64_t MOD[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull,
                                        0xb85045b68181585dull, 0x30644e72e131a029ull};
    static constexpr uint64_t R1[4] = {0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull,
                                       0x666ea36f7879462cull, 0x0e0a77c19a07df2full};
    static constexpr uint64_t R2[4] = {0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull,
                                       0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full};
    static constexpr uint64_t INV = 0x87d20782e4866389ull;  // -p^-1 mod 2^64
};
struct FrParams {
    static constexpr uint64_t MOD[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull,
                                        0xb85045b68181585dull, 0x30644e72e131a029ull};
    static constexpr uint64_t R1[4] = {0xac96341c4ffffffbull, 0x36fc76959f60cd29ull,
                                       0x666ea36f7879462eull, 0x0e0a77c19a07df2full};
    static constexpr uint64_t R2[4] = {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull,
                                       0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull};
    static constexpr uint64_t INV = 0xc2e1f593efffffffull;  // -r^-1 mod 2^64
};

// Field element held in Montgomery form, limbs least-significant first — the same
// in-memory form a Rust `&[Fr]` has, so it crosses the C ABI untouched.
template <class P>
struct Fp {
    uint64_t l[4];

    static Fp zero() { Fp r; r.l[0] = r.l[1] = r.l[2] = r.l[3] = 0; return r; }
    static Fp one() { Fp r; for (int i = 0; i < 4; i++) r.l[i] = P::R1[i]; return r; }
    static Fp from_raw(const uint64_t v[4]) {  // canonical integer (< modulus) -> Montgomery
        Fp a, r2;
        for (int i = 0; i < 4; i++) { a.l[i] = v[i]; r2.l[i] = P::R2[i]; }
        return a * r2;
    }
    static Fp from_u64(uint64_t v) { uint64_t t[4] = {v, 0, 0, 0}; return from_raw(t); }
    void to_raw(uint64_t out[4]) const {  // Montgomery -> canonical integer
        Fp o; o.l[0] = 1; o.l[1] = o.l[2] = o.l[3] = 0;
        Fp r = (*this) * o;
        for (int i = 0; i < 4; i++) out[i] = r.l[i];
    }
    bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
    bool operator==(const Fp& o) const {
        return l[0] == o.l[0] && l[1] == o.l[1] && l[2] == o.l[2] && l[3] == o.l[3];
    }
    bool operator!=(const Fp& o) const { return !(*this == o); }

    static bool geq_mod(const uint64_t v[4]) {
        for (int i = 3; i >= 0; i--) {
            if (v[i] > P::MOD[i]) return true;
            if (v[i] < P::MOD[i]) return false;
        }
        return true;
    }
    static void sub_mod(uint64_t v[4]) {
        u128 br = 0;
        for (int i = 0; i < 4; i++) {
            u128 d = (u128)v[i] - P::MOD[i] - br;
            v[i] = (uint64_t)d;
            br = (d >> 64) & 1;
        }
    }

    Fp operator+(const Fp& o) const {
        Fp r; u128 c = 0;
        for (int i = 0; i < 4; i++) { c += (u128)l[i] + o.l[i]; r.l[i] = (uint64_t)c; c >>= 64; }
        if (geq_mod(r.l)) sub_mod(r.l);  // both < 2^254 so no carry out of 256 bits
        return r;
    }
    Fp operator-(const Fp& o) const {
        Fp r; u128 br = 0;
        for (int i = 0; i < 4; i++) {
            u128 d = (u128)l[i] - o.l[i] - br;
            r.l[i] = (uint64_t)d; br = (d >> 64) & 1;
        }
        if (br) {
            u128 c = 0;
            for (int i = 0; i < 4; i++) { c += (u128)r.l[i] + P::MOD[i]; r.l[i] = (uint64_t)c; c >>= 64; }
        }
        return r;
    }
    Fp neg() const { return is_zero() ? *this : zero() - *this; }
    Fp dbl() const { return *this + *this; }

    // CIOS Montgomery product, 4 limbs of 64 bits.
    Fp operator*(const Fp& o) const {
        uint64_t t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; i++) {
            u128 c = 0;
            for (int j = 0; j < 4; j++) {
                c += (u128)l[j] * o.l[i] + t[j];
                t[j] = (uint64_t)c; c >>= 64;
            }
            c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
            uint64_t m = t[0] * P::INV;
            c = ((u128)m * P::MOD[0] + t[0]) >> 64;
            for (int j = 1; j < 4; j++) {
                c += (u128)m * P::MOD[j] + t[j];
                t[j - 1] = (uint64_t)c; c >>= 64;
            }
            c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
        }
        Fp r; for (int i = 0; i < 4; i++) r.l[i] = t[i];
        if (t[4] || geq_mod(r.l)) sub_mod(r.l);
        return r;
    }
    Fp sqr() const { return (*this) * (*this); }

    Fp pow(const uint64_t e[4]) const {
        Fp acc = one();
        for (int i = 255; i >= 0; i--) {
            acc = acc.sqr();
            if ((e[i >> 6] >> (i & 63)) & 1) acc = acc * (*this);
        }
        return acc;
    }
    Fp pow_u64(uint64_t e) const { uint64_t t[4] = {e, 0, 0, 0}; return pow(t); }
    Fp inv() const {  // Fermat; 0 -> 0
        uint64_t e[4]; for (int i = 0; i < 4; i++) e[i] = P::MOD[i];
        e[0] -= 2;  // low limb of both moduli is >= 2
        return pow(e);
    }
};

typedef Fp<FqParams> Fq;
typedef Fp<FrParams> Fr;

// Fr constants: 2-adicity 28, multiplicative generator 7.
static const uint64_t FR_ROOT_OF_UNITY_RAW[4] = {0xd34f1ed960c37c9cull, 0x3215cf6dd39329c8ull,
                                                 0x98865ea93dd31f74ull, 0x03ddb9f5166d18b7ull};
static const int FR_S = 28;

inline Fr fr_root_of_unity(int k) {  // omega_k = ROOT^(2^(28-k)), order 2^k
    Fr w = Fr::from_raw(FR_ROOT_OF_UNITY_RAW);
    for (int i = k; i < FR_S; i++) w = w.sqr();
    return w;
}

// ------------------------------------------------------------------ G1: y^2 = x^3 + 3
struct G1Affine {
    Fq x, y;  // identity encoded as (0, 0): (0,0) is not on the curve since b = 3
    bool is_identity() const { return x.is_zero() && y.is_zero(); }
    static G1Affine identity() { return G1Affine{Fq::zero(), Fq::zero()}; }
    static G1Affine generator() { return G1Affine{Fq::one(), Fq::from_u64(2)}; }
    G1Affine neg() const { return is_identity() ? *this : G1Affine{x, y.neg()}; }
    bool on_curve() const {
        if (is_identity()) return true;
        return y.sqr() == x.sqr() * x + Fq::from_u64(3);
    }
};

struct G1 {  // Jacobian (X/Z^2, Y/Z^3); identity <=> Z == 0
    Fq x, y, z;
    static G1 identity() { return G1{Fq::zero(), Fq::one(), Fq::zero()}; }
    static G1 from_affine(const G1Affine& a) {
        if (a.is_identity()) return identity();
        return G1{a.x, a.y, Fq::one()};
    }
    bool is_identity() const { return z.is_zero(); }
    G1 neg() const { return G1{x, y.neg(), z}; }

    G1 dbl() const {  // dbl-2009-l (a = 0)
        if (is_identity()) return *this;
        Fq a = x.sqr(), b = y.sqr(), c = b.sqr();
        Fq d = ((x + b).sqr() - a - c).dbl();
        Fq e = a.dbl() + a, f = e.sqr();
        Fq z3 = (z * y).dbl();
        Fq x3 = f - d.dbl();
        Fq c8 = c.dbl().dbl().dbl();
        Fq y3 = e * (d - x3) - c8;
        return G1{x3, y3, z3};
    }
    G1 add(const G1& o) const {  // add-2007-bl
        if (is_identity()) return o;
        if (o.is_identity()) return *this;
        Fq z1z1 = z.sqr(), z2z2 = o.z.sqr();
        Fq u1 = x * z2z2, u2 = o.x * z1z1;
        Fq s1 = y * z2z2 * o.z, s2 = o.y * z1z1 * z;
        if (u1 == u2) {
            if (s1 == s2) return dbl();
            return identity();
        }
        Fq h = u2 - u1, i = h.dbl().sqr(), j = h * i;
        Fq r = (s2 - s1).dbl(), v = u1 * i;
        Fq x3 = r.sqr() - j - v.dbl();
        Fq y3 = r * (v - x3) - (s1 * j).dbl();
        Fq z3 = ((z + o.z).sqr() - z1z1 - z2z2) * h;
        return G1{x3, y3, z3};
    }
    G1 add_mixed(const G1Affine& o) const {  // madd-2007-bl
        if (o.is_identity()) return *this;
        if (is_identity()) return from_affine(o);
        Fq z1z1 = z.sqr();
        Fq u2 = o.x * z1z1, s2 = o.y * z1z1 * z;
        if (x == u2) {
            if (y == s2) return dbl();
            return identity();
        }
        Fq h = u2 - x, hh = h.sqr(), i = hh.dbl().dbl(), j = h * i;
        Fq r = (s2 - y).dbl(), v = x * i;
        Fq x3 = r.sqr() - j - v.dbl();
        Fq y3 = r * (v - x3) - (y * j).dbl();
        Fq z3 = (z + h).sqr() - z1z1 - hh;
        return G1{x3, y3, z3};
    }
    G1Affine to_affine() const {
        if (is_identity()) return G1Affine::identity();
        Fq zi = z.inv(), zi2 = zi.sqr();
        return G1Affine{x * zi2, y * zi2 * zi};
    }
    G1 mul(const Fr& s) const {  // double-and-add over the canonical scalar
        uint64_t e[4]; s.to_raw(e);
        G1 acc = identity();
        for (int i = 255; i >= 0; i--) {
            acc = acc.dbl();
            if ((e[i >> 6] >> (i & 63)) & 1) acc = acc.add(*this);
        }
        return acc;
    }
};

}  // namespace orc
