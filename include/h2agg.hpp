// h2agg.hpp — the host side above the C ABI (include/h2agg.h) in C++17, header only.
//
// The reference (Trapdoor-Tech/halo2-aggregation) is Rust calling the `halo2` dependency; its hot path is reached
// through the functions below.  There is no Rust toolchain where this library is built, so this header is the
// compiled twin of the Rust wrappers in h2agg-shim/src/lib.rs — same names, same argument meaning, same error
// behaviour (a status other than H2A_OK becomes an `Error` carrying the code and h2a_last_error) — and
// examples/simple_example.cpp drives it the way examples/simple-example.rs drives the dependency.
//
//   dependency item (reached from the reference at)                                   here
//   arithmetic::best_multiexp(coeffs, bases)   examples/simple-example.rs:638-640     best_multiexp, Bases::msm
//   arithmetic::best_fft(a, omega, log_n)      domain used at src/verifier.rs:252,431 best_fft
//   poly::EvaluationDomain                     inside create_proof                    EvaluationDomain
//   Setup::new / Params::{read,write} / verifier_params   :589-590, :679-693          Params
//   keygen_vk / keygen_pk                      :593-594, :696-697                     Assembly, Circuit::set_keys
//   create_proof / verify_proof                :606-626, :702-728                     Circuit
//   Blake2bWrite / Blake2bRead, Challenge255   src/transcript.rs:58-133               Transcript
//   MultiopenChip::calc_witness, H fold        src/multiopen.rs:271-509, src/vanishing.rs:177-188   verify_accumulate, fold_h
//   ecc_chip.mul_var (witness cells)           src/multiopen.rs:393-492               mul_var_witness
//   point_to_scalars (68-bit limbs)            examples/simple-example.rs:535-548     point_to_scalars
//
// Field elements and points are the in-memory forms of the dependency's types: Fr = 4 x u64 little-endian limbs in
// Montgomery form, G1Affine = x || y (identity = 64 zero bytes).  There is no CPU path: every call that computes needs a
// B200 (h2a_init fails with H2A_ERR_NO_DEVICE otherwise).
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "h2agg.h"

namespace h2agg {

using Fr = std::array<uint8_t, 32>;        // bn256::Fr as it sits in memory
using G1Affine = std::array<uint8_t, 64>;  // bn256::G1Affine as it sits in memory

// What the reference gets as `halo2::plonk::Error` / a panic: the status code of the C ABI and the library's message.
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

// ---------------------------------------------------------------------------------------------------------------
// Host-side field values the way the reference writes them (`Fp::from(7)`, `-Fp::one()`, `Fr::DELTA`, the 68-bit limbs of
// a coordinate): a few Montgomery operations on 4 x u64 limbs so that a host program can build witness columns and public
// inputs without a device round trip.  Not a hot path.
namespace detail {
struct Modulus {
    uint64_t m[4], r2[4], inv;  // modulus, 2^512 mod m, -m^-1 mod 2^64
};
constexpr Modulus FR = {{0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
                        {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull},
                        0xc2e1f593efffffffull};
constexpr Modulus FQ = {{0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
                        {0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full},
                        0x87d20782e4866389ull};
inline bool geq(const uint64_t a[4], const uint64_t m[4]) {
    for (int i = 3; i >= 0; i--)
        if (a[i] != m[i]) return a[i] > m[i];
    return true;
}
inline void sub(uint64_t a[4], const uint64_t m[4]) {
    unsigned __int128 borrow = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned __int128 d = (unsigned __int128)a[i] - m[i] - borrow;
        a[i] = (uint64_t)d;
        borrow = (d >> 64) & 1;
    }
}
// Montgomery product a * b / 2^256 mod m (coarsely integrated operand scanning)
inline void mont_mul(const uint64_t a[4], const uint64_t b[4], const Modulus& f, uint64_t out[4]) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        unsigned __int128 carry = 0, cur;
        for (int j = 0; j < 4; j++) {
            cur = (unsigned __int128)a[j] * b[i] + t[j] + carry;
            t[j] = (uint64_t)cur;
            carry = cur >> 64;
        }
        cur = (unsigned __int128)t[4] + carry;
        t[4] = (uint64_t)cur;
        t[5] = (uint64_t)(cur >> 64);
        const uint64_t q = t[0] * f.inv;
        carry = ((unsigned __int128)q * f.m[0] + t[0]) >> 64;
        for (int j = 1; j < 4; j++) {
            cur = (unsigned __int128)q * f.m[j] + t[j] + carry;
            t[j - 1] = (uint64_t)cur;
            carry = cur >> 64;
        }
        cur = (unsigned __int128)t[4] + carry;
        t[3] = (uint64_t)cur;
        t[4] = t[5] + (uint64_t)(cur >> 64);
    }
    if (t[4] || geq(t, f.m)) sub(t, f.m);
    std::memcpy(out, t, 32);
}
}  // namespace detail

namespace fr {
inline Fr mul(const Fr& x, const Fr& y) {
    uint64_t a[4], b[4], o[4];
    std::memcpy(a, x.data(), 32);
    std::memcpy(b, y.data(), 32);
    detail::mont_mul(a, b, detail::FR, o);
    Fr out;
    std::memcpy(out.data(), o, 32);
    return out;
}
// a canonical integer below 2^128 as a field element: `Fr::from(v)`, `Fr::from_u128(v)`
inline Fr from_u128(uint64_t lo, uint64_t hi) {
    const uint64_t a[4] = {lo, hi, 0, 0};
    uint64_t o[4];
    detail::mont_mul(a, detail::FR.r2, detail::FR, o);
    Fr out;
    std::memcpy(out.data(), o, 32);
    return out;
}
inline Fr from_u64(uint64_t v) { return from_u128(v, 0); }
inline Fr zero() { return Fr{}; }
inline Fr one() { return from_u64(1); }
inline Fr add(const Fr& x, const Fr& y) {
    uint64_t a[4], b[4];
    std::memcpy(a, x.data(), 32);
    std::memcpy(b, y.data(), 32);
    unsigned __int128 carry = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned __int128 s = (unsigned __int128)a[i] + b[i] + carry;
        a[i] = (uint64_t)s;
        carry = s >> 64;
    }
    if (detail::geq(a, detail::FR.m)) detail::sub(a, detail::FR.m);  // r < 2^254: the sum of two elements has no carry out
    Fr out;
    std::memcpy(out.data(), a, 32);
    return out;
}
inline Fr neg(const Fr& x) {
    uint64_t a[4], o[4];
    std::memcpy(a, x.data(), 32);
    if (!(a[0] | a[1] | a[2] | a[3])) return x;
    std::memcpy(o, detail::FR.m, 32);
    detail::sub(o, a);
    Fr out;
    std::memcpy(out.data(), o, 32);
    return out;
}
inline Fr pow2k(Fr x, unsigned squarings) {  // x^(2^squarings)
    for (unsigned i = 0; i < squarings; i++) x = mul(x, x);
    return x;
}
// `Fr::DELTA` = 7^(2^28) (src/permutation.rs:259): the generator of the cosets the permutation columns are labelled with
inline Fr delta() { return pow2k(from_u64(7), 28); }
}  // namespace fr

// `point_to_scalars` of examples/simple-example.rs:535-548: x then y of an affine point, each as four 68-bit limbs (least
// significant first) of the canonical integer, every limb an Fr — the public inputs of the aggregation circuit (:668-672).
inline std::vector<Fr> point_to_scalars(const G1Affine& p) {
    std::vector<Fr> out;
    for (int c = 0; c < 2; c++) {
        uint64_t m[4], v[4];
        const uint64_t one[4] = {1, 0, 0, 0};
        std::memcpy(m, p.data() + 32 * c, 32);
        detail::mont_mul(m, one, detail::FQ, v);  // out of Montgomery form
        for (int l = 0; l < 4; l++) {
            const unsigned lo_bit = 68u * l, w = lo_bit / 64, s = lo_bit % 64;
            unsigned __int128 chunk = (unsigned __int128)v[w] >> s;
            if (w + 1 < 4) chunk |= (unsigned __int128)v[w + 1] << (64 - s);
            chunk &= (((unsigned __int128)1) << 68) - 1;
            out.push_back(fr::from_u128((uint64_t)chunk, (uint64_t)(chunk >> 64)));
        }
    }
    return out;
}

// The scalar `XorShiftRng::from_seed(seed)` yields first (examples/simple-example.rs:584-589: the KZG secret).  Host only.
inline Fr xorshift_scalar(const std::array<uint8_t, 16>& seed) {
    Fr out;
    const int rc = h2a_xorshift_scalar(seed.data(), out.data());
    if (rc != H2A_OK) throw Error(rc, "h2a_xorshift_scalar");
    return out;
}
// The transcript scalar of a verifying key (src/verifier.rs:341-358) from `format!("{:?}", vk.pinned())`: the vk_hash argument of
// Circuit::set_keys / Circuit::set_vk.  Host only.
inline Fr vk_hash(const std::string& pinned_debug) {
    Fr out;
    const int rc = h2a_vk_hash(reinterpret_cast<const uint8_t*>(pinned_debug.data()), pinned_debug.size(), out.data());
    if (rc != H2A_OK) throw Error(rc, "h2a_vk_hash");
    return out;
}
// `EvaluationDomain::get_omega` for 2^k rows.  Host only.
inline Fr root_of_unity(uint32_t k) {
    Fr out;
    const int rc = h2a_fr_root_of_unity(k, out.data());
    if (rc != H2A_OK) throw Error(rc, "h2a_fr_root_of_unity: k > 28");
    return out;
}
// Sum of affine points on the host (the combine step after an allgather of per-GPU partial MSMs).
inline G1Affine g1_sum(const std::vector<G1Affine>& points) {
    G1Affine out;
    const int rc = h2a_g1_sum(points.empty() ? nullptr : points[0].data(), points.size(), out.data());
    if (rc != H2A_OK) throw Error(rc, "h2a_g1_sum");
    return out;
}

// ---------------------------------------------------------------------------------------------------------------
// One GPU.  Calls on one context are serialised; one context per GPU (one process per GPU), distinct contexts from
// distinct threads.
class Context {
  public:
    explicit Context(int device = 0) {
        const int rc = h2a_init(&raw_, device);
        if (rc != H2A_OK) throw Error(rc, "h2a_init failed (no B200 / sm_100a device?)");
    }
    ~Context() {
        if (raw_) h2a_destroy(raw_);
    }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    h2a_ctx* raw() const { return raw_; }
    void check(int rc) const {
        if (rc != H2A_OK) throw Error(rc, h2a_last_error(raw_));
    }
    uint64_t launch_count() const { return h2a_launch_count(raw_); }

    // Several GPUs: rank 0 draws the two ids (comm_unique_ids) and hands them to every process over any channel.
    static std::array<uint8_t, 256> comm_unique_ids() {
        std::array<uint8_t, 256> ids{};
        for (int half = 0; half < 2; half++) {
            const int rc = h2a_comm_unique_id(ids.data() + 128 * half);
            if (rc != H2A_OK) throw Error(rc, "h2a_comm_unique_id failed (libnccl.so.2 not found?)");
        }
        return ids;
    }
    void comm_init(int rank, int world, const std::array<uint8_t, 256>& ids) { check(h2a_comm_init(raw_, rank, world, ids.data(), ids.data() + 128)); }
    int comm_world() const { return h2a_comm_world(raw_); }
    // raw bytes of every rank, in rank order
    std::vector<uint8_t> allgather(const std::vector<uint8_t>& send) {
        std::vector<uint8_t> out(send.size() * (size_t)comm_world());
        check(h2a_comm_allgather(raw_, send.data(), out.data(), send.size()));
        return out;
    }

  private:
    h2a_ctx* raw_ = nullptr;
};

// `Params.g` / `Params.g_lagrange` resident in HBM.
class Bases {
  public:
    Bases() = default;
    Bases(Context& ctx, h2a_bases* raw) : ctx_(&ctx), raw_(raw) {}
    // points: `&[G1Affine]`
    Bases(Context& ctx, const std::vector<G1Affine>& points) : ctx_(&ctx) {
        ctx.check(h2a_bases_upload(ctx.raw(), points.empty() ? nullptr : points[0].data(), points.size(), &raw_));
    }
    ~Bases() { reset(); }
    Bases(Bases&& o) noexcept : ctx_(o.ctx_), raw_(o.raw_) { o.raw_ = nullptr; }
    Bases& operator=(Bases&& o) noexcept {
        if (this != &o) {
            reset();
            ctx_ = o.ctx_;
            raw_ = o.raw_;
            o.raw_ = nullptr;
        }
        return *this;
    }
    Bases(const Bases&) = delete;
    Bases& operator=(const Bases&) = delete;

    size_t size() const { return h2a_bases_len(raw_); }
    const h2a_bases* raw() const { return raw_; }
    Context& ctx() const { return *ctx_; }
    // Window tables 2^(bits*w) * P_i, once per parameters (-1: the library's choice for single large MSMs; 17 suits provers).
    void precompute(int window_bits = -1) { ctx_->check(h2a_bases_precompute(ctx_->raw(), raw_, window_bits)); }
    // `best_multiexp(coeffs, &bases[..coeffs.len()])`
    G1Affine msm(const std::vector<Fr>& coeffs) const {
        G1Affine out;
        ctx_->check(h2a_msm_g1(ctx_->raw(), raw_, 0, coeffs.empty() ? nullptr : coeffs[0].data(), coeffs.size(), out.data()));
        return out;
    }
    // several polynomials over the same parameters in one call (the rounds of `create_proof`)
    std::vector<G1Affine> msm_batch(const std::vector<std::vector<Fr>>& columns) const {
        std::vector<const uint8_t*> ptrs;
        std::vector<size_t> lens;
        for (const auto& c : columns) {
            ptrs.push_back(c.empty() ? nullptr : c[0].data());
            lens.push_back(c.size());
        }
        std::vector<G1Affine> out(columns.size());
        if (!columns.empty()) ctx_->check(h2a_msm_g1_batch(ctx_->raw(), raw_, ptrs.data(), lens.data(), (int)columns.size(), out[0].data()));
        return out;
    }
    std::vector<G1Affine> download() const {
        std::vector<G1Affine> out(size());
        if (!out.empty()) ctx_->check(h2a_bases_download(ctx_->raw(), raw_, out[0].data()));
        return out;
    }

  private:
    void reset() {
        if (raw_) h2a_bases_free(ctx_->raw(), raw_);
        raw_ = nullptr;
    }
    Context* ctx_ = nullptr;
    h2a_bases* raw_ = nullptr;
};

// `halo2::arithmetic::best_multiexp` for bases that are not `Params` vectors (the verifier's sums).
inline G1Affine best_multiexp(Context& ctx, const std::vector<Fr>& coeffs, const std::vector<G1Affine>& bases) {
    if (coeffs.size() != bases.size()) throw Error(H2A_ERR_INVALID, "best_multiexp: coeffs and bases differ in length");
    G1Affine out;
    ctx.check(h2a_msm_g1_adhoc(ctx.raw(), bases.empty() ? nullptr : bases[0].data(), coeffs.empty() ? nullptr : coeffs[0].data(), coeffs.size(), out.data()));
    return out;
}

// `halo2::arithmetic::best_fft(a, omega, log_n)`: in place, natural order.
inline void best_fft(Context& ctx, std::vector<Fr>& a, const Fr& omega, uint32_t log_n) {
    if (a.size() != ((size_t)1 << log_n)) throw Error(H2A_ERR_INVALID, "best_fft: a.len() != 2^log_n");
    ctx.check(h2a_ntt(ctx.raw(), a[0].data(), log_n, omega.data(), 0, nullptr));
}

// The transforms of `halo2::poly::EvaluationDomain` (2^k rows, extended domain 2^extended_k, coset generator `zeta`).
class EvaluationDomain {
  public:
    // j = the circuit's degree, as in `EvaluationDomain::new(j, k)`; zeta = the dependency's coset generator
    EvaluationDomain(Context& ctx, uint32_t j, uint32_t k, const Fr& zeta) : ctx_(&ctx), k_(k), extended_k_(k), quotient_poly_degree_(j - 1), zeta_(zeta) {
        if (j < 2) throw Error(H2A_ERR_INVALID, "EvaluationDomain: degree below 2");
        while ((1ull << extended_k_) < (1ull << k) * (uint64_t)(j - 1)) extended_k_++;
        omega_ = root_of_unity(k);
    }
    uint32_t k() const { return k_; }
    uint32_t extended_k() const { return extended_k_; }
    Fr get_omega() const { return omega_; }
    size_t get_quotient_poly_degree() const { return quotient_poly_degree_; }
    void lagrange_to_coeff(std::vector<Fr>& a) const {
        if (a.size() != ((size_t)1 << k_)) throw Error(H2A_ERR_INVALID, "lagrange_to_coeff: wrong length");
        ctx_->check(h2a_ntt(ctx_->raw(), a[0].data(), k_, omega_.data(), 1, nullptr));
    }
    void coeff_to_lagrange(std::vector<Fr>& a) const {
        if (a.size() != ((size_t)1 << k_)) throw Error(H2A_ERR_INVALID, "coeff_to_lagrange: wrong length");
        ctx_->check(h2a_ntt(ctx_->raw(), a[0].data(), k_, omega_.data(), 0, nullptr));
    }
    std::vector<Fr> coeff_to_extended(const std::vector<Fr>& coeffs) const {
        if (coeffs.size() != ((size_t)1 << k_)) throw Error(H2A_ERR_INVALID, "coeff_to_extended: wrong length");
        std::vector<Fr> out((size_t)1 << extended_k_);
        ctx_->check(h2a_coeff_to_extended(ctx_->raw(), coeffs[0].data(), k_, extended_k_, zeta_.data(), out[0].data()));
        return out;
    }
    void extended_to_coeff(std::vector<Fr>& ext) const {
        if (ext.size() != ((size_t)1 << extended_k_)) throw Error(H2A_ERR_INVALID, "extended_to_coeff: wrong length");
        ctx_->check(h2a_extended_to_coeff(ctx_->raw(), ext[0].data(), extended_k_, zeta_.data()));
    }

  private:
    Context* ctx_;
    uint32_t k_, extended_k_;
    size_t quotient_poly_degree_;
    Fr omega_, zeta_;
};

// `Params<G1Affine>`: g and g_lagrange resident on the GPU.
class Params {
  public:
    uint32_t k = 0;
    Bases g, g_lagrange;
    bool has_trailer = false;
    std::array<uint8_t, 128> trailer{};  // the opaque tail of a parameter file ([s]G2 for the pairing check; the caller's)

    // `Setup::<Bn256>::new(k, rng)` once `rng` has produced the secret (examples/simple-example.rs:589, :687)
    static Params setup(Context& ctx, uint32_t k, const Fr& secret) {
        h2a_bases *g = nullptr, *gl = nullptr;
        ctx.check(h2a_kzg_setup(ctx.raw(), k, secret.data(), &g, &gl));
        Params p;
        p.k = k;
        p.g = Bases(ctx, g);
        p.g_lagrange = Bases(ctx, gl);
        return p;
    }
    // `Params::read` (:681-684); the byte format is the library's (csrc/params.cu)
    static Params read(Context& ctx, const std::string& path) {
        h2a_bases *g = nullptr, *gl = nullptr;
        Params p;
        int has = 0;
        ctx.check(h2a_params_read(ctx.raw(), path.c_str(), &p.k, &g, &gl, p.trailer.data(), &has));
        p.has_trailer = has != 0;
        p.g = Bases(ctx, g);
        p.g_lagrange = Bases(ctx, gl);
        return p;
    }
    // `Params::write` (:686-690)
    void write(const std::string& path, bool compressed = false) const {
        g.ctx().check(h2a_params_write(g.ctx().raw(), path.c_str(), k, g.raw(), g_lagrange.raw(), compressed ? 1 : 0, has_trailer ? trailer.data() : nullptr));
    }
    // `Setup::verifier_params(&params, public_inputs_size)` (:590, :693): the bases `commit_lagrange(public_inputs)` needs.
    // Shares the memory of g_lagrange, which must outlive it.
    Bases verifier_params(size_t public_inputs_size) const {
        h2a_bases* raw = nullptr;
        g.ctx().check(h2a_params_verifier_view(g.ctx().raw(), g_lagrange.raw(), public_inputs_size, &raw));
        return Bases(g.ctx(), raw);
    }
    size_t get_n() const { return (size_t)1 << k; }
    G1Affine commit(const std::vector<Fr>& poly_coeffs) const { return g.msm(poly_coeffs); }
    G1Affine commit_lagrange(const std::vector<Fr>& poly_evals) const { return g_lagrange.msm(poly_evals); }
};

// ---------------------------------------------------------------------------------------------------------------
// What `VerifierChip::_verify_proof` reads from the verifying key (src/verifier.rs:286-311), as the word stream of
// csrc/plonk_shape.hpp.  Expressions are postfix programs of (op, argument) pairs over the query lists.
enum Op : uint32_t { OP_CONST = 0, OP_ADVICE, OP_FIXED, OP_INSTANCE, OP_NEG, OP_ADD, OP_MUL, OP_SCALE };
enum ColumnType : uint32_t { ADVICE = 0, FIXED = 1, INSTANCE = 2 };
struct Query {
    uint32_t column;
    int32_t rotation;
};
using Program = std::vector<std::pair<uint32_t, uint32_t>>;
struct LookupArgument {
    std::vector<Program> inputs, tables;
};
struct PermutationColumn {
    uint32_t type, column, query_index;  // query_index: the column's rotation-0 query (get_any_query_index)
};
struct CircuitShape {
    uint32_t k = 0, blinding_factors = 0, degree = 0, num_instance = 0, num_advice = 0, num_fixed = 0;
    std::vector<Query> advice_queries, fixed_queries, instance_queries;
    std::vector<Program> gates;
    std::vector<Fr> constants;
    std::vector<LookupArgument> lookups;
    std::vector<PermutationColumn> permutation_columns;

    size_t n() const { return (size_t)1 << k; }
    size_t usable_rows() const { return n() - (blinding_factors + 1); }
    std::vector<uint32_t> words() const {
        std::vector<uint32_t> w = {0x48324153u, k, blinding_factors, degree, num_instance, num_advice, num_fixed};
        for (const auto* qs : {&advice_queries, &fixed_queries, &instance_queries}) {
            w.push_back((uint32_t)qs->size());
            for (const Query& q : *qs) {
                w.push_back(q.column);
                w.push_back((uint32_t)q.rotation);
            }
        }
        auto prog = [&w](const Program& p) {
            w.push_back((uint32_t)p.size());
            for (const auto& oa : p) {
                w.push_back(oa.first);
                w.push_back(oa.second);
            }
        };
        w.push_back((uint32_t)gates.size());
        for (const Program& g : gates) prog(g);
        w.push_back((uint32_t)constants.size());
        w.push_back((uint32_t)lookups.size());
        for (const LookupArgument& l : lookups) {
            w.push_back((uint32_t)l.inputs.size());
            for (const Program& p : l.inputs) prog(p);
            w.push_back((uint32_t)l.tables.size());
            for (const Program& p : l.tables) prog(p);
        }
        w.push_back((uint32_t)permutation_columns.size());
        for (const PermutationColumn& c : permutation_columns) {
            w.push_back(c.type);
            w.push_back(c.column);
            w.push_back(c.query_index);
        }
        return w;
    }
};

// The copy-constraint bookkeeping of `keygen_vk` / `keygen_pk` (:593-594): `copy` joins the cycles of two cells of the
// permutation columns (host), `sigmas` evaluates the sigma columns on the device.
class Assembly {
  public:
    Assembly(uint32_t n_cols, uint32_t k) : n_cols_(n_cols), k_(k) {
        const int rc = h2a_assembly_new(n_cols, k, &raw_);
        if (rc != H2A_OK) throw Error(rc, "h2a_assembly_new");
    }
    ~Assembly() {
        if (raw_) h2a_assembly_free(raw_);
    }
    Assembly(const Assembly&) = delete;
    Assembly& operator=(const Assembly&) = delete;
    void copy(uint32_t col_a, uint32_t row_a, uint32_t col_b, uint32_t row_b) {
        const int rc = h2a_assembly_copy(raw_, col_a, row_a, col_b, row_b);
        if (rc != H2A_OK) throw Error(rc, "h2a_assembly_copy: cell outside the permutation columns");
    }
    // next cell of every cell in its cycle: col' * 2^k + row'
    std::vector<uint32_t> mapping() const {
        std::vector<uint32_t> out((size_t)n_cols_ << k_);
        const int rc = h2a_assembly_mapping(raw_, out.data());
        if (rc != H2A_OK) throw Error(rc, "h2a_assembly_mapping");
        return out;
    }
    // n_cols columns of 2^k elements, column-major: the `sigmas` argument of Circuit::set_keys
    std::vector<Fr> sigmas(Context& ctx, const Fr& omega, const Fr& delta) const {
        std::vector<Fr> out((size_t)n_cols_ << k_);
        ctx.check(h2a_assembly_sigmas(ctx.raw(), raw_, omega.data(), delta.data(), out[0].data()));
        return out;
    }

  private:
    h2a_assembly* raw_ = nullptr;
    uint32_t n_cols_, k_;
};

// A circuit with its verifying key (and proving key once set_keys ran): `create_proof` / `verify_proof` for it.
class Circuit {
  public:
    Circuit(Context& ctx, const CircuitShape& shape) : ctx_(&ctx), shape_(shape) {
        const std::vector<uint32_t> w = shape.words();
        ctx.check(h2a_circuit_create(ctx.raw(), w.data(), w.size(), shape.constants.empty() ? nullptr : shape.constants[0].data(), shape.constants.size(), &raw_));
    }
    ~Circuit() {
        if (raw_) h2a_circuit_free(ctx_->raw(), raw_);
    }
    Circuit(const Circuit&) = delete;
    Circuit& operator=(const Circuit&) = delete;
    const CircuitShape& shape() const { return shape_; }

    // `keygen_pk` output: fixed columns and permutation columns (column-major, n elements each) are committed and kept on the
    // device; vk_hash is the transcript scalar of the verifying key (src/verifier.rs:341-358), zeta the coset generator.
    void set_keys(const Params& params, const std::vector<Fr>& fixed_values, const std::vector<Fr>& sigmas, const Fr& vk_hash, const Fr& zeta) {
        if (fixed_values.size() != shape_.n() * shape_.num_fixed || sigmas.size() != shape_.n() * shape_.permutation_columns.size())
            throw Error(H2A_ERR_INVALID, "set_keys: fixed / sigma columns do not hold n elements per column");
        ctx_->check(h2a_circuit_set_keys(ctx_->raw(), raw_, params.g.raw(), params.g_lagrange.raw(), fixed_values.empty() ? nullptr : fixed_values[0].data(),
                                         sigmas.empty() ? nullptr : sigmas[0].data(), vk_hash.data(), zeta.data()));
    }
    // verifier only: the commitments of the verifying key
    void set_vk(const std::vector<G1Affine>& fixed_commitments, const std::vector<G1Affine>& sigma_commitments, const Fr& vk_hash) {
        if (fixed_commitments.size() != shape_.num_fixed || sigma_commitments.size() != shape_.permutation_columns.size())
            throw Error(H2A_ERR_INVALID, "set_vk: wrong number of commitments");
        ctx_->check(h2a_circuit_set_vk(ctx_->raw(), raw_, fixed_commitments.empty() ? nullptr : fixed_commitments[0].data(),
                                       sigma_commitments.empty() ? nullptr : sigma_commitments[0].data(), vk_hash.data()));
    }
    // (fixed commitments, sigma commitments) of the verifying key
    std::pair<std::vector<G1Affine>, std::vector<G1Affine>> get_vk() const {
        std::vector<G1Affine> f(shape_.num_fixed), s(shape_.permutation_columns.size());
        ctx_->check(h2a_circuit_get_vk(ctx_->raw(), raw_, f.empty() ? nullptr : f[0].data(), s.empty() ? nullptr : s[0].data()));
        return {f, s};
    }
    // one proof over the ranks of the context's communicator (Context::comm_init); world = 1 switches it off
    void distribute(int rank, int world) { ctx_->check(h2a_circuit_set_distribution(ctx_->raw(), raw_, rank, world, nullptr, nullptr)); }
    size_t blinds_len() const { return h2a_blinds_len(raw_); }
    size_t proof_len() const { return h2a_proof_len(raw_); }

    struct Proof {
        std::vector<uint8_t> bytes;                  // `transcript.finalize()`
        std::vector<G1Affine> instance_commitments;  // what the verifier recomputes with commit_lagrange(public_inputs)
    };
    // `create_proof(&params, &pk, &[circuit], &[&[&public_inputs]], &mut transcript)`: columns column-major, advice already
    // blinded, `blinds` in the order include/h2agg.h gives for h2a_create_proof.
    Proof create_proof(const std::vector<Fr>& instance_cols, const std::vector<Fr>& advice_cols, const std::vector<Fr>& blinds) {
        if (blinds.size() != blinds_len()) throw Error(H2A_ERR_INVALID, "create_proof: wrong number of blinding scalars");
        if (instance_cols.size() != shape_.n() * shape_.num_instance || advice_cols.size() != shape_.n() * shape_.num_advice)
            throw Error(H2A_ERR_INVALID, "create_proof: instance / advice columns do not hold n elements per column");
        Proof p;
        p.bytes.resize(proof_len());
        p.instance_commitments.resize(shape_.num_instance);
        size_t len = 0;
        ctx_->check(h2a_create_proof(ctx_->raw(), raw_, instance_cols.empty() ? nullptr : instance_cols[0].data(), advice_cols.empty() ? nullptr : advice_cols[0].data(),
                                     blinds[0].data(), p.bytes.data(), p.bytes.size(), &len, p.instance_commitments.empty() ? nullptr : p.instance_commitments[0].data()));
        p.bytes.resize(len);
        return p;
    }
    // `verify_proof(&params_verifier, vk, instances, &mut transcript)` up to the pairing: [e, f, w, zw]
    // (examples/simple-example.rs:620, :668-671; src/verifier.rs:739-742).  Malformed bytes: Error with code H2A_ERR_PROOF.
    std::array<G1Affine, 4> verify_proof(const std::vector<G1Affine>& instance_commitments, const std::vector<uint8_t>& proof) const {
        if (instance_commitments.size() != shape_.num_instance) throw Error(H2A_ERR_INVALID, "verify_proof: wrong number of instance commitments");
        std::array<G1Affine, 4> out;
        ctx_->check(h2a_verify_proof(ctx_->raw(), raw_, instance_commitments.empty() ? nullptr : instance_commitments[0].data(), proof.data(), proof.size(), out[0].data()));
        return out;
    }
    // a batch of independent proofs of this circuit in one launch (BASELINE config 5); instance commitments proof-major
    std::vector<std::array<G1Affine, 4>> verify_proof_batch(const std::vector<G1Affine>& instance_commitments, const std::vector<std::vector<uint8_t>>& proofs) const {
        if (instance_commitments.size() != proofs.size() * shape_.num_instance) throw Error(H2A_ERR_INVALID, "verify_proof_batch: wrong number of instance commitments");
        std::vector<const uint8_t*> ptrs;
        std::vector<size_t> lens;
        for (const auto& p : proofs) {
            ptrs.push_back(p.data());
            lens.push_back(p.size());
        }
        std::vector<std::array<G1Affine, 4>> out(proofs.size());
        if (!proofs.empty())
            ctx_->check(h2a_verify_proof_batch(ctx_->raw(), raw_, proofs.size(), instance_commitments.empty() ? nullptr : instance_commitments[0].data(), ptrs.data(), lens.data(),
                                               out[0][0].data()));
        return out;
    }

  private:
    Context* ctx_;
    CircuitShape shape_;
    h2a_circuit* raw_ = nullptr;
};

// The multi-open accumulation alone (`MultiopenChip::calc_witness`, src/multiopen.rs:271-509) for callers that keep their
// own transcript replay: [e, f, w, zw].
inline std::array<G1Affine, 4> verify_accumulate(Context& ctx, const std::vector<G1Affine>& commitments, const std::vector<int32_t>& rotations, const std::vector<Fr>& evals,
                                                  const std::vector<G1Affine>& ws, const Fr& x, const Fr& u, const Fr& v, const Fr& omega, const G1Affine& g1) {
    if (commitments.size() != rotations.size() || rotations.size() != evals.size()) throw Error(H2A_ERR_INVALID, "verify_accumulate: query lists differ in length");
    std::array<G1Affine, 4> out;
    ctx.check(h2a_verify_accumulate(ctx.raw(), commitments.empty() ? nullptr : commitments[0].data(), rotations.data(), evals.empty() ? nullptr : evals[0].data(), rotations.size(),
                                    ws.empty() ? nullptr : ws[0].data(), ws.size(), x.data(), u.data(), v.data(), omega.data(), g1.data(), out[0].data()));
    return out;
}

// H = sum_i (x^n)^i h_i (src/vanishing.rs:177-188)
inline G1Affine fold_h(Context& ctx, const std::vector<G1Affine>& h_pieces, const Fr& xn) {
    G1Affine out;
    ctx.check(h2a_fold_h(ctx.raw(), h_pieces.empty() ? nullptr : h_pieces[0].data(), h_pieces.size(), xn.data(), out.data()));
    return out;
}

// Witness cells of the non-native `ecc_chip.mul_var(region, point, scalar, offset)` of the aggregation circuit
// (src/multiopen.rs:393-492, src/vanishing.rs:181-187) for a batch of (point, scalar) pairs: results[i] = scalars[i] * points[i]
// and mul_var_witness_len() Fr cells per pair (layout: csrc/mulvar.cu).  `aux` is the auxiliary point the incomplete additions
// start from.  status[i] != 0 marks a pair the incomplete formulas cannot witness (`ok` is then false), as an unsatisfiable
// circuit would be reported; other failures throw.
struct MulVarWitness {
    bool ok = true;
    std::vector<G1Affine> results;
    std::vector<Fr> cells;
    std::vector<uint32_t> status;
};
inline size_t mul_var_witness_len() { return h2a_mulvar_witness_len(); }
inline MulVarWitness mul_var_witness(Context& ctx, const std::vector<G1Affine>& points, const std::vector<Fr>& scalars, const G1Affine& aux, bool want_cells = true) {
    if (points.size() != scalars.size()) throw Error(H2A_ERR_INVALID, "mul_var_witness: points and scalars differ in length");
    MulVarWitness w;
    const size_t m = points.size();
    w.results.resize(m);
    w.status.assign(m, 0);
    if (want_cells) w.cells.resize(m * mul_var_witness_len());
    if (m == 0) return w;
    const int rc = h2a_mulvar_witness(ctx.raw(), points[0].data(), scalars[0].data(), m, aux.data(), w.results[0].data(), want_cells ? w.cells[0].data() : nullptr, w.status.data());
    bool flagged = false;
    for (uint32_t st : w.status) flagged = flagged || st != 0;
    if (rc == H2A_ERR_INVALID && flagged) {
        w.ok = false;
        return w;
    }
    ctx.check(rc);
    return w;
}

// `Blake2bWrite<_, _, Challenge255<_>>` / `Blake2bRead` (src/transcript.rs:58,72,105-107,122-124).  Host only.
class Transcript {
  public:
    Transcript() : raw_(h2a_transcript_new()) {
        if (!raw_) throw Error(H2A_ERR_OOM, "h2a_transcript_new");
    }
    ~Transcript() { h2a_transcript_free(raw_); }
    Transcript(const Transcript&) = delete;
    Transcript& operator=(const Transcript&) = delete;
    // false for the identity, which the transcript refuses (as the dependency's `common_point` does)
    bool common_point(const G1Affine& p) { return h2a_transcript_common_point(raw_, p.data()) == H2A_OK; }
    void common_scalar(const Fr& s) { h2a_transcript_common_scalar(raw_, s.data()); }
    Fr squeeze_challenge() {
        Fr out;
        h2a_transcript_squeeze_challenge(raw_, out.data());
        return out;
    }

  private:
    h2a_transcript* raw_;
};

}  // namespace h2agg
