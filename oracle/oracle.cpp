// oracle/oracle.cpp — CPU restatement of the hot path's algorithms, exported with a C ABI so
// tests can drive it through ctypes.
//
// TEST INFRASTRUCTURE ONLY (see bn254.hpp header).  PARITY UNPINNED (see bn254.hpp header).
//
// What is restated here, and which reference lines consume it:
//   * best_multiexp  — halo2 `arithmetic::best_multiexp`; reached from
//                      examples/simple-example.rs:638-640 (commit_lagrange) and every
//                      commitment inside create_proof (:606-613, :702-709).
//   * best_fft       — halo2 `arithmetic::best_fft` + EvaluationDomain ops;
//                      the reference touches the domain at src/verifier.rs:252,431.
//   * Blake2b transcript — src/transcript.rs:58,72,105-107,122-124.
//   * GWC accumulation   — src/multiopen.rs:19-45 (set ordering), :271-509 (Horner chains).
//   * H fold             — src/vanishing.rs:177-188.
#include "bn254.hpp"

#include <algorithm>
#include <cmath>
#include <map>
#include <thread>
#include <vector>

using namespace orc;

namespace {

// ------------------------------------------------------------------ synthetic inputs
// Counter-based generator shared (by specification, not by code) with the device library:
// see include/h2agg.h `h2a_gen_scalars` / `h2a_gen_bases`.
inline uint64_t mix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
inline uint64_t draw254(uint64_t seed, uint64_t stream, uint64_t i, uint64_t attempt, uint64_t out[4]) {
    uint64_t h = mix64(seed + 0x100000001b3ull * stream);
    h = mix64(h ^ i);
    h = mix64(h + attempt);
    for (int j = 0; j < 4; j++) out[j] = mix64(h + (uint64_t)j + 1);
    out[3] &= 0x3fffffffffffffffull;
    return mix64(h + 5);  // spare bits (sign choice)
}

template <class F>
void parallel_for(size_t n, int threads, F f) {
    if (threads <= 1 || n < 2) { f(0, n, 0); return; }
    std::vector<std::thread> ts;
    size_t chunk = (n + threads - 1) / threads;
    for (int t = 0; t < threads; t++) {
        size_t lo = std::min(n, chunk * t), hi = std::min(n, lo + chunk);
        if (lo >= hi) break;
        ts.emplace_back([=] { f(lo, hi, t); });
    }
    for (auto& t : ts) t.join();
}

inline Fr load_fr(const uint8_t* p) { Fr r; memcpy(r.l, p, 32); return r; }
inline void store_fr(uint8_t* p, const Fr& v) { memcpy(p, v.l, 32); }
inline G1Affine load_affine(const uint8_t* p) { G1Affine a; memcpy(a.x.l, p, 32); memcpy(a.y.l, p + 32, 32); return a; }
inline void store_affine(uint8_t* p, const G1Affine& a) { memcpy(p, a.x.l, 32); memcpy(p + 32, a.y.l, 32); }

// ------------------------------------------------------------------ best_multiexp restatement
// Bucket states as upstream: None / Affine (first insert stays affine) / Projective.
struct Bucket {
    int state = 0;  // 0 none, 1 affine, 2 projective
    G1Affine a;
    G1 p;
    void add_assign(const G1Affine& o) {
        if (state == 0) { a = o; state = 1; }
        else if (state == 1) { p = G1::from_affine(a).add_mixed(o); state = 2; }
        else p = p.add_mixed(o);
    }
    G1 add_to(const G1& acc) const {
        if (state == 0) return acc;
        if (state == 1) return acc.add_mixed(a);
        return acc.add(p);
    }
};

inline uint64_t get_at(size_t segment, size_t c, const uint8_t bytes[32]) {
    size_t skip_bits = segment * c;
    size_t skip_bytes = skip_bits / 8;
    if (skip_bytes >= 32) return 0;
    uint8_t v[8] = {0};
    for (size_t i = 0; i < 8 && skip_bytes + i < 32; i++) v[i] = bytes[skip_bytes + i];
    uint64_t tmp; memcpy(&tmp, v, 8);
    tmp >>= skip_bits - skip_bytes * 8;
    return tmp % (1ull << c);
}

G1 multiexp_serial(const uint8_t* scalars_mont, const uint8_t* bases, size_t n) {
    std::vector<uint8_t> repr(n * 32);
    for (size_t i = 0; i < n; i++) {  // to_repr(): canonical little-endian
        uint64_t raw[4]; load_fr(scalars_mont + 32 * i).to_raw(raw);
        memcpy(&repr[32 * i], raw, 32);
    }
    size_t c;
    if (n < 4) c = 1; else if (n < 32) c = 3; else c = (size_t)std::ceil(std::log((double)n));
    size_t segments = 256 / c + 1;
    G1 acc = G1::identity();
    std::vector<Bucket> buckets((1u << c) - 1);
    for (size_t seg = segments; seg-- > 0;) {
        for (size_t k = 0; k < c; k++) acc = acc.dbl();
        for (auto& b : buckets) b.state = 0;
        for (size_t i = 0; i < n; i++) {
            uint64_t d = get_at(seg, c, &repr[32 * i]);
            if (d != 0) buckets[d - 1].add_assign(load_affine(bases + 64 * i));
        }
        G1 running = G1::identity();
        for (size_t k = buckets.size(); k-- > 0;) {
            running = buckets[k].add_to(running);
            acc = acc.add(running);
        }
    }
    return acc;
}

G1 best_multiexp(const uint8_t* scalars, const uint8_t* bases, size_t n, int threads) {
    if (threads < 1) threads = 1;
    if (n > (size_t)threads && threads > 1) {
        size_t chunk = n / threads;  // upstream: chunks(chunk) -> possibly threads+1 pieces
        size_t pieces = (n + chunk - 1) / chunk;
        std::vector<G1> partial(pieces, G1::identity());
        std::vector<std::thread> ts;
        for (size_t t = 0; t < pieces; t++) {
            size_t lo = t * chunk, hi = std::min(n, lo + chunk);
            ts.emplace_back([=, &partial] { partial[t] = multiexp_serial(scalars + 32 * lo, bases + 64 * lo, hi - lo); });
        }
        for (auto& t : ts) t.join();
        G1 acc = G1::identity();
        for (auto& p : partial) acc = acc.add(p);
        return acc;
    }
    return multiexp_serial(scalars, bases, n);
}

// ------------------------------------------------------------------ best_fft restatement
inline uint32_t bitrev(uint32_t v, int bits) {
    uint32_t r = 0;
    for (int i = 0; i < bits; i++) { r = (r << 1) | (v & 1); v >>= 1; }
    return r;
}

void serial_fft(Fr* a, size_t n, const Fr& omega, int log_n) {
    for (size_t k = 0; k < n; k++) {
        size_t rk = bitrev((uint32_t)k, log_n);
        if (k < rk) std::swap(a[k], a[rk]);
    }
    size_t m = 1;
    for (int s = 0; s < log_n; s++) {
        Fr w_m = omega.pow_u64(n / (2 * m));
        for (size_t k = 0; k < n; k += 2 * m) {
            Fr w = Fr::one();
            for (size_t j = 0; j < m; j++) {
                Fr t = a[k + j + m] * w;
                a[k + j + m] = a[k + j] - t;
                a[k + j] = a[k + j] + t;
                w = w * w_m;
            }
        }
        m *= 2;
    }
}

void best_fft(Fr* a, const Fr& omega, int log_n, int threads) {
    int log_threads = 0;
    while ((2 << log_threads) <= threads) log_threads++;
    size_t n = (size_t)1 << log_n;
    if (log_n <= log_threads || log_threads == 0) { serial_fft(a, n, omega, log_n); return; }
    // bellman-style split into 2^log_threads sub-transforms of size n / 2^log_threads
    size_t T = (size_t)1 << log_threads;
    int log_new_n = log_n - log_threads;
    size_t new_n = (size_t)1 << log_new_n;
    std::vector<std::vector<Fr>> tmp(T, std::vector<Fr>(new_n, Fr::zero()));
    Fr new_omega = omega.pow_u64(T);
    std::vector<std::thread> ts;
    for (size_t j = 0; j < T; j++) {
        ts.emplace_back([&, j] {
            Fr omega_j = omega.pow_u64(j);
            Fr omega_step = omega.pow_u64(j << log_new_n);
            Fr elt = Fr::one();
            for (size_t i = 0; i < new_n; i++) {
                for (size_t s = 0; s < T; s++) {
                    size_t idx = (i + (s << log_new_n)) & (n - 1);
                    tmp[j][i] = tmp[j][i] + a[idx] * elt;
                    elt = elt * omega_step;
                }
                elt = elt * omega_j;
            }
            serial_fft(tmp[j].data(), new_n, new_omega, log_new_n);
        });
    }
    for (auto& t : ts) t.join();
    parallel_for(n, threads, [&](size_t lo, size_t hi, int) {
        for (size_t idx = lo; idx < hi; idx++) a[idx] = tmp[idx & (T - 1)][idx >> log_threads];
    });
}

// ------------------------------------------------------------------ Blake2b (RFC 7693)
struct Blake2b {
    uint64_t h[8];
    uint64_t t0 = 0, t1 = 0;
    uint8_t buf[128];
    size_t buflen = 0;
    size_t outlen;

    static constexpr uint64_t IV[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull,
                                       0xa54ff53a5f1d36f1ull, 0x510e527fade682d1ull, 0x9b05688c2b3e6c1full,
                                       0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
    Blake2b(size_t out_len, const uint8_t personal[16]) : outlen(out_len) {
        uint8_t param[64] = {0};
        param[0] = (uint8_t)out_len; param[2] = 1; param[3] = 1;  // digest len, fanout, depth
        if (personal) memcpy(param + 48, personal, 16);
        for (int i = 0; i < 8; i++) { uint64_t w; memcpy(&w, param + 8 * i, 8); h[i] = IV[i] ^ w; }
    }
    static inline uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
    void compress(const uint8_t block[128], bool last) {
        static const uint8_t S[12][16] = {
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
            {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
            {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
            {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
            {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
        uint64_t m[16], v[16];
        memcpy(m, block, 128);
        for (int i = 0; i < 8; i++) { v[i] = h[i]; v[i + 8] = IV[i]; }
        v[12] ^= t0; v[13] ^= t1;
        if (last) v[14] = ~v[14];
        auto G = [&](int a, int b, int c, int d, uint64_t x, uint64_t y) {
            v[a] = v[a] + v[b] + x; v[d] = rotr(v[d] ^ v[a], 32);
            v[c] = v[c] + v[d];     v[b] = rotr(v[b] ^ v[c], 24);
            v[a] = v[a] + v[b] + y; v[d] = rotr(v[d] ^ v[a], 16);
            v[c] = v[c] + v[d];     v[b] = rotr(v[b] ^ v[c], 63);
        };
        for (int r = 0; r < 12; r++) {
            const uint8_t* s = S[r];
            G(0, 4, 8, 12, m[s[0]], m[s[1]]);   G(1, 5, 9, 13, m[s[2]], m[s[3]]);
            G(2, 6, 10, 14, m[s[4]], m[s[5]]);  G(3, 7, 11, 15, m[s[6]], m[s[7]]);
            G(0, 5, 10, 15, m[s[8]], m[s[9]]);  G(1, 6, 11, 12, m[s[10]], m[s[11]]);
            G(2, 7, 8, 13, m[s[12]], m[s[13]]); G(3, 4, 9, 14, m[s[14]], m[s[15]]);
        }
        for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
    }
    void update(const uint8_t* in, size_t len) {
        while (len) {
            if (buflen == 128) {  // keep the last block for finalisation
                t0 += 128; if (t0 < 128) t1++;
                compress(buf, false); buflen = 0;
            }
            size_t take = std::min(len, 128 - buflen);
            memcpy(buf + buflen, in, take);
            buflen += take; in += take; len -= take;
        }
    }
    void finalize(uint8_t* out) const {  // const: the transcript squeezes from a clone
        Blake2b c = *this;
        c.t0 += c.buflen; if (c.t0 < c.buflen) c.t1++;
        memset(c.buf + c.buflen, 0, 128 - c.buflen);
        c.compress(c.buf, true);
        memcpy(out, c.h, outlen);
    }
};
constexpr uint64_t Blake2b::IV[8];

// Fr::from_bytes_wide: 512-bit little-endian integer mod r, result in Montgomery form.
Fr fr_from_bytes_wide(const uint8_t b[64]) {
    uint64_t lo[4], hi[4];
    memcpy(lo, b, 32); memcpy(hi, b + 32, 32);
    // value = lo + hi*2^256.  Montgomery trick: mont(x) = x*R;  lo*R = mul(lo, R2); hi*2^256*R = mul(hi, R3).
    Fr l, h, r2, r3;
    for (int i = 0; i < 4; i++) { l.l[i] = lo[i]; h.l[i] = hi[i]; r2.l[i] = FrParams::R2[i]; }
    r3 = r2 * r2;  // R^3 in plain terms: mont-mul(R2,R2) = R^2*R^2/R = R^3
    // mul() tolerates unreduced inputs < 2^256 because the CIOS bound only needs one operand < modulus.
    return l * r2 + h * r3;
}

struct Transcript {  // Blake2bWrite / Blake2bRead state with Challenge255
    Blake2b st;
    Transcript() : st(64, (const uint8_t*)"Halo2-Transcript") {}
    void common_point(const G1Affine& p) {
        uint8_t buf[65]; buf[0] = 1;
        uint64_t raw[4];
        p.x.to_raw(raw); memcpy(buf + 1, raw, 32);
        p.y.to_raw(raw); memcpy(buf + 33, raw, 32);
        st.update(buf, 65);
    }
    void common_scalar(const Fr& s) {
        uint8_t buf[33]; buf[0] = 2;
        uint64_t raw[4]; s.to_raw(raw); memcpy(buf + 1, raw, 32);
        st.update(buf, 33);
    }
    Fr squeeze() {
        uint8_t z = 0; st.update(&z, 1);
        uint8_t wide[64]; st.finalize(wide);
        return fr_from_bytes_wide(wide);
    }
};

}  // namespace

// ====================================================================== C ABI (ctypes)
// ---- non-native mul_var witness cells (row f4): the same cells as oracle/mulvar.py, compiled, so that the benchmark has a CPU
// figure that is not an interpreter's.  Written the way sequential witness code computes them — one field inversion per affine
// formula, 512-bit product and exact quotient per record — cell order as documented in csrc/mulvar.cu; checked cell for cell against
// the Python big-integer statement in tests/test_mulvar_oracle.py.  halo2wrong (Cargo.toml:10) is not in the tree: PARITY UNPINNED.
namespace mulvar {
constexpr int BITS = 254, REC = 22, STEP = 7 * REC + 16, FINAL = 3 * REC + 8, LEN = BITS + BITS * STEP + FINAL;
typedef unsigned __int128 u128;
struct W4 { uint64_t w[4]; };   // up to 256 bits
inline W4 w4(uint64_t lo = 0) { W4 r{{lo, 0, 0, 0}}; return r; }
inline void add(W4& a, const W4& b) { u128 c = 0; for (int i = 0; i < 4; i++) { c += (u128)a.w[i] + b.w[i]; a.w[i] = (uint64_t)c; c >>= 64; } }
inline void sub(W4& a, const W4& b) { u128 br = 0; for (int i = 0; i < 4; i++) { u128 d = (u128)a.w[i] - b.w[i] - br; a.w[i] = (uint64_t)d; br = (d >> 64) & 1; } }
inline W4 shl68(const W4& a) { W4 r = w4(); for (int i = 0; i < 3; i++) { r.w[i + 1] |= a.w[i] << 4; if (i + 2 < 4) r.w[i + 2] |= a.w[i] >> 60; } return r; }
inline W4 shr136(const W4& a) { W4 r = w4(); r.w[0] = (a.w[2] >> 8) | (a.w[3] << 56); r.w[1] = a.w[3] >> 8; return r; }
struct Limb { uint64_t lo, hi; };   // 68 bits: hi < 16
inline Limb limb_of(const uint64_t v[4], int i) {
    const int bit = 68 * i, w = bit >> 6, sh = bit & 63;
    const uint64_t x0 = v[w], x1 = w + 1 < 4 ? v[w + 1] : 0, x2 = w + 2 < 4 ? v[w + 2] : 0;
    Limb l;
    l.lo = sh ? (x0 >> sh) | (x1 << (64 - sh)) : x0;
    l.hi = (sh ? (x1 >> sh) | (x2 << (64 - sh)) : x1) & 0xf;
    return l;
}
inline W4 mul68(const Limb& a, const Limb& b) {   // 136 bits
    W4 r = w4();
    u128 p = (u128)a.lo * b.lo;
    r.w[0] = (uint64_t)p; r.w[1] = (uint64_t)(p >> 64);
    u128 mid = (u128)a.lo * b.hi + (u128)b.lo * a.hi;            // < 2^69
    W4 m = w4(); m.w[1] = (uint64_t)mid; m.w[2] = (uint64_t)(mid >> 64);
    add(r, m);
    W4 h = w4(); h.w[2] = a.hi * b.hi;
    add(r, h);
    return r;
}
inline Fr cell(const W4& v) { return Fr::from_raw(v.w); }          // values are far below r
inline Fr cell(const Limb& l) { uint64_t t[4] = {l.lo, l.hi, 0, 0}; return Fr::from_raw(t); }
// p' = 2^272 - p as four 68-bit limbs; p^-1 mod 2^256
static const Limb NEG_P[4] = {{0xc3df73e9278302b9ull, 0x2}, {0x2687e956e978e357ull, 0xa}, {0xd647afba497e7ea7ull, 0xf}, {0xfffcf9bb18d1ece5ull, 0xf}};
static const uint64_t P_INV[4] = {0x782df87d1b799c77ull, 0x6121829ae1359536ull, 0x2750342fe7cc257full, 0x0a85dd486e777394ull};
inline void mul_low(const uint64_t a[4], const uint64_t b[4], uint64_t out[4]) {
    uint64_t r[4] = {0, 0, 0, 0};
    for (int i = 0; i < 4; i++) { u128 c = 0; for (int j = 0; i + j < 4; j++) { c += (u128)a[i] * b[j] + r[i + j]; r[i + j] = (uint64_t)c; c >>= 64; } }
    for (int i = 0; i < 4; i++) out[i] = r[i];
}
// cells of a * b = q p + r; returns r (Montgomery)
inline Fq record(const Fq& a_m, const Fq& b_m, Fr* out) {
    const Fq r_m = a_m * b_m;
    uint64_t a[4], b[4], r[4], lo[4], q[4];
    a_m.to_raw(a); b_m.to_raw(b); r_m.to_raw(r);
    mul_low(a, b, lo);
    { u128 br = 0; for (int i = 0; i < 4; i++) { u128 d = (u128)lo[i] - r[i] - br; lo[i] = (uint64_t)d; br = (d >> 64) & 1; } }
    mul_low(lo, P_INV, q);
    Limb al[4], bl[4], ql[4], rl[4];
    for (int i = 0; i < 4; i++) { al[i] = limb_of(a, i); bl[i] = limb_of(b, i); ql[i] = limb_of(q, i); rl[i] = limb_of(r, i); }
    for (int i = 0; i < 4; i++) { out[i] = cell(al[i]); out[4 + i] = cell(bl[i]); out[8 + i] = cell(ql[i]); out[12 + i] = cell(rl[i]); }
    W4 t[4];
    for (int k = 0; k < 4; k++) {
        t[k] = w4();
        for (int i = 0; i <= k; i++) { add(t[k], mul68(al[i], bl[k - i])); add(t[k], mul68(ql[i], NEG_P[k - i])); }
        out[16 + k] = cell(t[k]);
    }
    W4 v = w4();
    for (int half = 0; half < 2; half++) {
        W4 u = t[2 * half];
        add(u, shl68(t[2 * half + 1]));
        add(u, v);
        W4 rr = w4(), r1 = w4();
        rr.w[0] = rl[2 * half].lo; rr.w[1] = rl[2 * half].hi;
        r1.w[0] = rl[2 * half + 1].lo; r1.w[1] = rl[2 * half + 1].hi;
        add(rr, shl68(r1));
        sub(u, rr);
        v = shr136(u);
        out[20 + half] = cell(v);
    }
    return r_m;
}
inline void point_cells(const G1Affine& p, Fr* out) {
    uint64_t x[4], y[4];
    p.x.to_raw(x); p.y.to_raw(y);
    for (int i = 0; i < 4; i++) { out[i] = cell(limb_of(x, i)); out[4 + i] = cell(limb_of(y, i)); }
}
inline bool add_cells(const G1Affine& a, const G1Affine& b, Fr* out, G1Affine& res) {
    const Fq dx = b.x - a.x;
    if (dx.is_zero()) return false;
    const Fq lam = (b.y - a.y) * dx.inv();
    record(lam, dx, out);
    const Fq l2 = record(lam, lam, out + REC);
    res.x = l2 - a.x - b.x;
    res.y = record(lam, a.x - res.x, out + 2 * REC) - a.y;
    return true;
}
inline void double_cells(const G1Affine& a, Fr* out, G1Affine& res) {
    const Fq xx = record(a.x, a.x, out);
    const Fq y2 = a.y.dbl();
    const Fq lam = (xx.dbl() + xx) * y2.inv();
    record(lam, y2, out + REC);
    const Fq l2 = record(lam, lam, out + 2 * REC);
    res.x = l2 - a.x.dbl();
    res.y = record(lam, a.x - res.x, out + 3 * REC) - a.y;
}
// one mul_var; returns the status (0 ok, 1 + step, 0xffffffff identity input)
inline uint32_t witness(const G1Affine& P, const Fr& s, const G1Affine& aux, const G1Affine& corr, Fr* cells, G1Affine& Q) {
    uint64_t e[4];
    s.to_raw(e);
    auto bit = [&](int b) { return (e[b >> 6] >> (b & 63)) & 1; };
    for (int b = 0; b < BITS; b++) cells[b] = bit(b) ? Fr::one() : Fr::zero();
    Q = G1Affine::identity();
    if (P.is_identity()) return 0xffffffffu;
    G1Affine acc = aux;
    for (int step = 0; step < BITS; step++) {
        Fr* base = cells + BITS + (size_t)step * STEP;
        G1Affine D, T;
        double_cells(acc, base, D);
        if (!add_cells(D, P, base + 4 * REC, T)) return 1 + step;
        point_cells(D, base + 7 * REC);
        acc = bit(BITS - 1 - step) ? T : D;
        point_cells(acc, base + 7 * REC + 8);
    }
    Fr* base = cells + BITS + (size_t)BITS * STEP;
    if (!add_cells(acc, corr, base, Q)) { Q = G1Affine::identity(); return 1 + BITS; }
    point_cells(Q, base + 3 * REC);
    return 0;
}
}  // namespace mulvar


extern "C" {

int orc_hw_threads() { return (int)std::thread::hardware_concurrency(); }

// ---- field element-wise (device field-arithmetic parity) ; op: 0 add 1 sub 2 mul 3 sqr 4 inv 5 neg
void orc_field_op(int field, int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        if (field == 0) {
            Fq x, y, r; memcpy(x.l, a + 32 * i, 32); if (b) memcpy(y.l, b + 32 * i, 32);
            switch (op) { case 0: r = x + y; break; case 1: r = x - y; break; case 2: r = x * y; break;
                          case 3: r = x.sqr(); break; case 4: r = x.inv(); break; default: r = x.neg(); }
            memcpy(out + 32 * i, r.l, 32);
        } else {
            Fr x, y, r; memcpy(x.l, a + 32 * i, 32); if (b) memcpy(y.l, b + 32 * i, 32);
            switch (op) { case 0: r = x + y; break; case 1: r = x - y; break; case 2: r = x * y; break;
                          case 3: r = x.sqr(); break; case 4: r = x.inv(); break; default: r = x.neg(); }
            memcpy(out + 32 * i, r.l, 32);
        }
    }
}
void orc_to_mont(int field, const uint8_t* canonical, uint8_t* out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        uint64_t raw[4]; memcpy(raw, canonical + 32 * i, 32);
        if (field == 0) { Fq r = Fq::from_raw(raw); memcpy(out + 32 * i, r.l, 32); }
        else { Fr r = Fr::from_raw(raw); memcpy(out + 32 * i, r.l, 32); }
    }
}
void orc_from_mont(int field, const uint8_t* mont, uint8_t* out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        uint64_t raw[4];
        if (field == 0) { Fq r; memcpy(r.l, mont + 32 * i, 32); r.to_raw(raw); }
        else { Fr r; memcpy(r.l, mont + 32 * i, 32); r.to_raw(raw); }
        memcpy(out + 32 * i, raw, 32);
    }
}
void orc_fr_root_of_unity(int k, uint8_t out[32]) { store_fr(out, fr_root_of_unity(k)); }

// ---- synthetic inputs (same specification as the device generators)
void orc_gen_scalars(uint64_t seed, size_t first, size_t n, uint8_t* out, int threads) {
    parallel_for(n, threads, [&](size_t lo, size_t hi, int) {
        for (size_t i = lo; i < hi; i++) {
            uint64_t v[4];
            for (uint64_t att = 0;; att++) {
                draw254(seed, 1, first + i, att, v);
                if (!Fr::geq_mod(v)) break;
            }
            memcpy(out + 32 * i, v, 32);  // the drawn value IS the in-memory (Montgomery) form
        }
    });
}
void orc_gen_bases(uint64_t seed, size_t first, size_t n, uint8_t* out, int threads) {
    // (p+1)/4
    static const uint64_t E[4] = {0x4f082305b61f3f52ull, 0x65e05aa45a1c72a3ull, 0x6e14116da0605617ull, 0x0c19139cb84c680aull};
    parallel_for(n, threads, [&](size_t lo, size_t hi, int) {
        Fq three = Fq::from_u64(3);
        for (size_t i = lo; i < hi; i++) {
            for (uint64_t att = 0;; att++) {
                uint64_t v[4];
                uint64_t spare = draw254(seed, 2, first + i, att, v);
                if (Fq::geq_mod(v)) continue;
                Fq x = Fq::from_raw(v);
                Fq rhs = x.sqr() * x + three;
                Fq y = rhs.pow(E);
                if (y.sqr() != rhs) continue;
                if (spare & 1) y = y.neg();
                store_affine(out + 64 * i, G1Affine{x, y});
                break;
            }
        }
    });
}

// ---- G1
int orc_g1_on_curve(const uint8_t* pts, size_t n) {
    for (size_t i = 0; i < n; i++) if (!load_affine(pts + 64 * i).on_curve()) return 0;
    return 1;
}
void orc_g1_add(const uint8_t a[64], const uint8_t b[64], uint8_t out[64]) {
    store_affine(out, G1::from_affine(load_affine(a)).add_mixed(load_affine(b)).to_affine());
}
void orc_g1_mul(const uint8_t a[64], const uint8_t s[32], uint8_t out[64]) {
    store_affine(out, G1::from_affine(load_affine(a)).mul(load_fr(s)).to_affine());
}
// best_multiexp restatement; out = canonical affine (x||y Montgomery, identity = zeros)
void orc_msm(const uint8_t* bases, const uint8_t* scalars, size_t n, int threads, uint8_t out[64]) {
    store_affine(out, best_multiexp(scalars, bases, n, threads).to_affine());
}
// independent check of the above: plain sum of double-and-add products
void orc_msm_naive(const uint8_t* bases, const uint8_t* scalars, size_t n, int threads, uint8_t out[64]) {
    if (threads < 1) threads = 1;
    std::vector<G1> part(threads, G1::identity());
    parallel_for(n, threads, [&](size_t lo, size_t hi, int t) {
        G1 acc = G1::identity();
        for (size_t i = lo; i < hi; i++) acc = acc.add(G1::from_affine(load_affine(bases + 64 * i)).mul(load_fr(scalars + 32 * i)));
        part[t] = acc;
    });
    G1 acc = G1::identity();
    for (auto& p : part) acc = acc.add(p);
    store_affine(out, acc.to_affine());
}

// ---- NTT (in place, natural order in and out)
void orc_fft(uint8_t* a, int log_n, const uint8_t omega[32], int threads) {
    best_fft((Fr*)a, load_fr(omega), log_n, threads);
}
void orc_ifft(uint8_t* a, int log_n, const uint8_t omega_inv[32], int threads) {  // EvaluationDomain::ifft
    size_t n = (size_t)1 << log_n;
    Fr* v = (Fr*)a;
    best_fft(v, load_fr(omega_inv), log_n, threads);
    Fr ninv = Fr::from_u64(n).inv();
    parallel_for(n, threads, [&](size_t lo, size_t hi, int) { for (size_t i = lo; i < hi; i++) v[i] = v[i] * ninv; });
}
// coeff_to_extended with an explicit coset shift g: out[i] = in[i]*g^i (i<n), zero-pad to 2^ext_k, FFT(ext omega)
void orc_coeff_to_extended(const uint8_t* coeffs, int k, int ext_k, const uint8_t shift[32], uint8_t* out, int threads) {
    size_t n = (size_t)1 << k, m = (size_t)1 << ext_k;
    Fr* o = (Fr*)out; const Fr* c = (const Fr*)coeffs;
    Fr g = load_fr(shift);
    parallel_for(m, threads, [&](size_t lo, size_t hi, int) {
        Fr p = g.pow_u64(lo);
        for (size_t i = lo; i < hi; i++) { o[i] = (i < n) ? c[i] * p : Fr::zero(); p = p * g; }
    });
    best_fft(o, fr_root_of_unity(ext_k), ext_k, threads);
}
// extended_to_coeff: iFFT(ext omega^-1, 1/m), then out[i] *= g^-i ; caller truncates
void orc_extended_to_coeff(uint8_t* ext, int ext_k, const uint8_t shift[32], int threads) {
    size_t m = (size_t)1 << ext_k;
    Fr* v = (Fr*)ext;
    Fr winv = fr_root_of_unity(ext_k).inv();
    best_fft(v, winv, ext_k, threads);
    Fr minv = Fr::from_u64(m).inv();
    Fr ginv = load_fr(shift).inv();
    parallel_for(m, threads, [&](size_t lo, size_t hi, int) {
        Fr p = ginv.pow_u64(lo) * minv;
        for (size_t i = lo; i < hi; i++) { v[i] = v[i] * p; p = p * ginv; }
    });
}

// ---- Blake2b + transcript
void orc_blake2b(const uint8_t* msg, size_t len, const uint8_t* personal16, size_t outlen, uint8_t* out) {
    Blake2b b(outlen, personal16); b.update(msg, len); b.finalize(out);
}
void orc_fr_from_bytes_wide(const uint8_t in[64], uint8_t out[32]) { store_fr(out, fr_from_bytes_wide(in)); }

void* orc_transcript_new() { return new Transcript(); }
void orc_transcript_free(void* t) { delete (Transcript*)t; }
void orc_transcript_common_point(void* t, const uint8_t p[64]) { ((Transcript*)t)->common_point(load_affine(p)); }
void orc_transcript_common_scalar(void* t, const uint8_t s[32]) { ((Transcript*)t)->common_scalar(load_fr(s)); }
void orc_transcript_squeeze(void* t, uint8_t out[32]) { store_fr(out, ((Transcript*)t)->squeeze()); }

// ---- GWC accumulation, following MultiopenChip::calc_witness step by step
// (src/multiopen.rs:271-509).  commitments: nq*64, rotations: nq int32, evals: nq*32, ws: S*64.
// Returns 0, or -1 when the number of W points differs from the number of rotation sets.
int orc_gwc_accumulate(const uint8_t* commitments, const int32_t* rotations, const uint8_t* evals, size_t nq,
                       const uint8_t* ws, size_t n_ws, const uint8_t x_[32], const uint8_t u_[32],
                       const uint8_t v_[32], const uint8_t omega_[32], const uint8_t g1_[64],
                       uint8_t out_efwzw[4 * 64]) {
    std::map<int32_t, std::vector<size_t>> sets;  // BTreeMap<Rotation, Vec<Q>>  (:25)
    for (size_t i = 0; i < nq; i++) sets[rotations[i]].push_back(i);
    if (sets.size() != n_ws) return -1;
    Fr x = load_fr(x_), u = load_fr(u_), v = load_fr(v_), omega = load_fr(omega_), omega_inv = omega.inv();
    std::vector<G1> Ws, ZWs, Fs;
    Fr eval_multi = Fr::zero();
    size_t si = 0;
    for (auto& kv : sets) {
        int32_t r = kv.first;
        Fr omega_eval = r >= 0 ? omega.pow_u64((uint64_t)r) : omega_inv.pow_u64((uint64_t)(-(int64_t)r));  // :348-359
        Fr z = omega_eval * x;                                                                           // :385-390
        G1 wi = G1::from_affine(load_affine(ws + 64 * si++));
        Ws.push_back(wi);
        ZWs.push_back(wi.mul(z));                                                                        // :393
        eval_multi = eval_multi * u;                                                                     // :406-409
        G1 cb = G1::from_affine(load_affine(commitments + 64 * kv.second[0]));
        Fr eb = load_fr(evals + 32 * kv.second[0]);
        for (size_t j = 1; j < kv.second.size(); j++) {                                                  // :416-462
            size_t q = kv.second[j];
            cb = cb.mul(v).add_mixed(load_affine(commitments + 64 * q));
            eb = eb * v + load_fr(evals + 32 * q);
        }
        Fs.push_back(cb);
        eval_multi = eval_multi + eb;                                                                    // :467-469
    }
    auto horner = [&](std::vector<G1>& pts) {                                                            // :471-488
        G1 acc = pts[0];
        for (size_t i = 1; i < pts.size(); i++) acc = acc.mul(u).add(pts[i]);
        return acc;
    };
    G1 w = horner(Ws), zw = horner(ZWs), f = horner(Fs);
    G1 e = G1::from_affine(load_affine(g1_)).mul(eval_multi.neg());                                       // :490-492
    store_affine(out_efwzw + 0, e.to_affine());
    store_affine(out_efwzw + 64, f.to_affine());
    store_affine(out_efwzw + 128, w.to_affine());
    store_affine(out_efwzw + 192, zw.to_affine());
    return 0;
}

// H = sum_i (x^n)^i h_i  as the reference folds it (src/vanishing.rs:177-188):
size_t orc_mulvar_witness_len() { return mulvar::LEN; }
// m independent mul_var over the host threads; cells may be NULL (results and status only)
void orc_mulvar_witness(const uint8_t* points, const uint8_t* scalars, size_t m, const uint8_t aux_[64], uint8_t* results, uint8_t* cells,
                        uint32_t* status, int threads) {
    const G1Affine aux = load_affine(aux_);
    G1 c = G1::from_affine(aux);
    for (int i = 0; i < mulvar::BITS; i++) c = c.dbl();
    const G1Affine corr = c.to_affine().neg();
    parallel_for(m, threads, [&](size_t lo, size_t hi, int) {
        std::vector<Fr> tmp(cells ? 0 : mulvar::LEN);
        for (size_t i = lo; i < hi; i++) {
            Fr* out = cells ? (Fr*)(cells + 32ull * mulvar::LEN * i) : tmp.data();
            G1Affine q;
            status[i] = mulvar::witness(load_affine(points + 64 * i), load_fr(scalars + 32 * i), aux, corr, out, q);
            store_affine(results + 64 * i, q);
        }
    });
}

// H starts at h_0; each later piece is multiplied by a running power of x^n and added.
void orc_fold_h(const uint8_t* h_pieces, size_t n_pieces, const uint8_t xn_[32], uint8_t out[64]) {
    Fr xn = load_fr(xn_), xn_power = xn;
    G1 acc = G1::from_affine(load_affine(h_pieces));
    for (size_t i = 1; i < n_pieces; i++) {
        G1 term = G1::from_affine(load_affine(h_pieces + 64 * i)).mul(xn_power);  // :181-184
        xn_power = xn_power * xn;                                                  // :185
        acc = acc.add(term);                                                       // :186
    }
    store_affine(out, acc.to_affine());
}

}  // extern "C"
