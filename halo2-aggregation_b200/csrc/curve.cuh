// curve.cuh — BN254 G1 (y^2 = x^3 + 3) on the device: affine inputs, XYZZ accumulators.
//
// XYZZ (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2) keeps the mixed addition at 8M + 2S with no inversion,
// which is what the bucket accumulation of the Pippenger MSM spends its time in.  The value an MSM
// returns is a unique group element (halo2 `best_multiexp`, reached from
// examples/simple-example.rs:638-640), so the coordinate system is free; results leave the library
// in canonical affine form.
#pragma once
#include "field.cuh"

namespace h2a {

struct Affine {  // identity <=> (0, 0)   ((0,0) is not on the curve: b = 3)
    Fq x, y;
    __device__ __forceinline__ bool is_identity() const { return x.is_zero() && y.is_zero(); }
    __device__ __forceinline__ static Affine load(const void* p) {
        Affine a;
        a.x = Fq::load(p);
        a.y = Fq::load((const uint8_t*)p + 32);
        return a;
    }
    __device__ __forceinline__ static Affine load_gather(const void* p) {   // see Fp::load_gather
        Affine a;
        a.x = Fq::load_gather(p);
        a.y = Fq::load_gather((const uint8_t*)p + 32);
        return a;
    }
    __device__ __forceinline__ void store(void* p) const {
        x.store(p);
        y.store((uint8_t*)p + 32);
    }
};

struct XYZZ {  // identity <=> zz == 0
    Fq x, y, zz, zzz;

    __device__ __forceinline__ static XYZZ identity() {
        XYZZ r;
        r.x = Fq::zero(); r.y = Fq::zero(); r.zz = Fq::zero(); r.zzz = Fq::zero();
        return r;
    }
    __device__ __forceinline__ bool is_identity() const { return zz.is_zero(); }
    __device__ __forceinline__ static XYZZ from_affine(const Affine& a) {
        XYZZ r;
        if (a.is_identity()) return identity();
        r.x = a.x; r.y = a.y; r.zz = Fq::one(); r.zzz = Fq::one();
        return r;
    }
    __device__ __forceinline__ static XYZZ load(const void* p) {
        XYZZ r;
        const uint8_t* b = (const uint8_t*)p;
        r.x = Fq::load(b); r.y = Fq::load(b + 32); r.zz = Fq::load(b + 64); r.zzz = Fq::load(b + 96);
        return r;
    }
    __device__ __forceinline__ void store(void* p) const {
        uint8_t* b = (uint8_t*)p;
        x.store(b); y.store(b + 32); zz.store(b + 64); zzz.store(b + 96);
    }

    // 2*(affine point), mdbl-2008-s-1 with a = 0.  a must not be the identity.
    __device__ __noinline__ static XYZZ dbl_affine(const Affine& a) {
        XYZZ r;
        Fq u = a.y.dbl();
        Fq v = u.sqr();
        Fq w = u * v;
        Fq s = a.x * v;
        Fq xx = a.x.sqr();
        Fq m = xx.dbl() + xx;
        r.x = m.sqr() - s.dbl();
        r.y = m * (s - r.x) - w * a.y;
        r.zz = v;
        r.zzz = w;
        return r;
    }
    // dbl-2008-s-1 with a = 0
    __device__ __noinline__ XYZZ dbl() const {
        if (is_identity()) return *this;
        XYZZ r;
        Fq u = y.dbl();
        Fq v = u.sqr();
        Fq w = u * v;
        Fq s = x * v;
        Fq xx = x.sqr();
        Fq m = xx.dbl() + xx;
        r.x = m.sqr() - s.dbl();
        r.y = m * (s - r.x) - w * y;
        r.zz = v * zz;
        r.zzz = w * zzz;
        return r;
    }

    // this += (a.x, neg ? -a.y : a.y): madd-2008-s (8M + 2S); a must not be the identity.
    __device__ __forceinline__ void add_affine(const Affine& a_in, bool neg) {
        Affine a = a_in;
        if (neg) a.y = a.y.neg();
        if (is_identity()) {
            x = a.x; y = a.y; zz = Fq::one(); zzz = Fq::one();
            return;
        }
        Fq u2 = a.x * zz;
        Fq s2 = a.y * zzz;
        Fq p = u2 - x;
        Fq r = s2 - y;
        if (p.is_zero()) {  // same x: doubling or cancellation (rare; warp-divergent path)
            if (r.is_zero()) *this = dbl_affine(a);
            else *this = identity();
            return;
        }
        Fq pp = p.sqr();
        Fq ppp = p * pp;
        Fq q = x * pp;
        Fq x3 = r.sqr() - ppp - q.dbl();
        y = r * (q - x3) - y * ppp;
        x = x3;
        zz = zz * pp;
        zzz = zzz * ppp;
    }

    // this += o: add-2008-s (12M + 2S)
    __device__ __forceinline__ void add(const XYZZ& o) {
        if (o.is_identity()) return;
        if (is_identity()) { *this = o; return; }
        Fq u1 = x * o.zz;
        Fq u2 = o.x * zz;
        Fq s1 = y * o.zzz;
        Fq s2 = o.y * zzz;
        Fq p = u2 - u1;
        Fq r = s2 - s1;
        if (p.is_zero()) {
            if (r.is_zero()) *this = dbl();
            else *this = identity();
            return;
        }
        Fq pp = p.sqr();
        Fq ppp = p * pp;
        Fq q = u1 * pp;
        Fq x3 = r.sqr() - ppp - q.dbl();
        y = r * (q - x3) - s1 * ppp;
        x = x3;
        zz = zz * o.zz * pp;
        zzz = zzz * o.zzz * ppp;
    }
};

}  // namespace h2a
