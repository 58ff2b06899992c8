// plonk_verify.cu — the verifier-side transcript glue: replays what `VerifierChip::_verify_proof`
// (src/verifier.rs:286-762) does natively — wire order, challenges, l_0/l_last/l_blind, gate / permutation /
// lookup expressions, expected h(x), the query list — and hands ALL sums of the GWC accumulation
// (src/multiopen.rs:271-509), with H flattened into its pieces (src/vanishing.rs:177-188), to one launch of
// the small-MSM kernel.  Output: (e, f, w, zw), the `[G1Affine; 4]` of examples/simple-example.rs:620,668-671.
#include <algorithm>
#include <cstring>
#include <string>
#include <thread>

#include "host_glue.hpp"
#include "plonk.hpp"

int h2a_small_msm(h2a_ctx* ctx, const uint8_t* bases, const uint8_t* scalars, const uint32_t* sum_offsets, size_t n_sums,
                  uint8_t* out_affine);

namespace {
using namespace h2a_host;
using namespace h2a_plonk;

static const uint64_t FQ_SQRT_EXP[4] = {0x4f082305b61f3f52ull, 0x65e05aa45a1c72a3ull, 0x6e14116da0605617ull,
                                        0x0c19139cb84c680aull};  // (p+1)/4

// 32-byte compressed point: x little-endian, bit 255 = parity of y, identity = zeros
bool decompress_point(const uint8_t* b, PointA& out) {
    uint8_t t[32];
    memcpy(t, b, 32);
    bool all_zero = true;
    for (int i = 0; i < 32; i++) all_zero &= t[i] == 0;
    if (all_zero) { out = PointA{fq_zero(), fq_zero()}; return true; }
    unsigned sign = t[31] >> 7;
    t[31] &= 0x7f;
    uint64_t raw[4];
    memcpy(raw, t, 32);
    if (ge_p(raw, MOD_Q)) return false;
    Fq x = fq_from_raw(raw);
    Fq rhs = sqr(x) * x + Fq{el_from_u64(3, MOD_Q)};
    Fq y = Fq{el_pow(rhs.e, FQ_SQRT_EXP, MOD_Q)};
    if (!(sqr(y) == rhs)) return false;
    uint64_t yr[4];
    fq_to_raw(y, yr);
    if ((yr[0] & 1) != sign) y = neg(y);
    out = PointA{x, y};
    return true;
}

struct Reader {
    const uint8_t* p;
    size_t len, pos = 0;
    bool ok = true;
    bool point(PointA& out) {
        if (pos + 32 > len || !decompress_point(p + pos, out)) return ok = false;
        pos += 32;
        return true;
    }
    bool scalar(Fr& out) {
        if (pos + 32 > len) return ok = false;
        uint64_t raw[4];
        memcpy(raw, p + pos, 32);
        if (ge_p(raw, MOD_R)) return ok = false;
        out = fr_from_raw(raw);
        pos += 32;
        return true;
    }
};

// error text of one proof's replay (the proofs of a batch are replayed by several host threads)
struct ErrSink {
    std::string err;
};

// One proof -> its (e, f, w, zw) term lists appended to `tl`.  Host only; touches nothing shared but `c` (read-only).
int collect_terms(ErrSink* ctx, const h2a_circuit* c, const uint8_t* inst_comms, const uint8_t* proof, size_t len,
                  const uint8_t g1[64], h2a_glue::TermList& tl) {
    const Shape& s = c->shape;
    h2a_glue::Transcript tr;
    Reader rd{proof, len};
    auto rpoint = [&](PointA& p) { return rd.point(p) && tr.common_point(p); };
    auto rscalar = [&](Fr& v) { if (!rd.scalar(v)) return false; tr.common_scalar(v); return true; };
#define NEED(x) do { if (!(x)) H2A_FAIL(ctx, H2A_ERR_PROOF, "verify: malformed proof at byte %zu", rd.pos); } while (0)

    tr.common_scalar(fr_load(c->vk_hash));                                        // src/verifier.rs:341-358
    std::vector<PointA> inst(s.n_instance);
    for (uint32_t i = 0; i < s.n_instance; i++) {                                 // :360-363
        inst[i] = affine_load(inst_comms + 64 * i);
        if (!tr.common_point(inst[i])) H2A_FAIL(ctx, H2A_ERR_INVALID, "verify: instance commitment %u is the identity", i);
    }
    std::vector<PointA> adv(s.n_advice);
    for (auto& p : adv) NEED(rpoint(p));                                          // :365-376
    Fr theta = tr.squeeze();                                                      // :378
    std::vector<PointA> lk_a(s.lookups.size()), lk_s(s.lookups.size()), lk_z(s.lookups.size());
    for (size_t i = 0; i < s.lookups.size(); i++) { NEED(rpoint(lk_a[i])); NEED(rpoint(lk_s[i])); }   // :380-387
    Fr beta = tr.squeeze(), gamma = tr.squeeze();                                 // :390,393
    std::vector<PointA> pz(s.n_chunks);
    for (auto& p : pz) NEED(rpoint(p));                                           // :402-409
    for (auto& p : lk_z) NEED(rpoint(p));                                         // :411-417
    PointA random_comm;
    NEED(rpoint(random_comm));                                                    // :419-421
    Fr y = tr.squeeze();                                                          // :423
    std::vector<PointA> h(s.qdeg);
    for (auto& p : h) NEED(rpoint(p));                                            // :427-434
    Fr x = tr.squeeze();                                                          // :436
    std::vector<Fr> ie(s.iq.size()), ae(s.aq.size()), fe(s.fq.size());
    for (auto& v : ie) NEED(rscalar(v));                                          // :438-475
    for (auto& v : ae) NEED(rscalar(v));
    for (auto& v : fe) NEED(rscalar(v));
    Fr random_eval;
    NEED(rscalar(random_eval));                                                   // src/vanishing.rs:108-134
    std::vector<Fr> se(s.perm.size());
    for (auto& v : se) NEED(rscalar(v));                                          // src/permutation.rs:140-168
    struct PSet { Fr z, zn, zl; bool has_last; };
    std::vector<PSet> ps(s.n_chunks);
    for (uint32_t i = 0; i < s.n_chunks; i++) {                                   // src/permutation.rs:81-138
        ps[i].has_last = i + 1 < s.n_chunks;
        NEED(rd.scalar(ps[i].z));
        NEED(rd.scalar(ps[i].zn));
        if (ps[i].has_last) NEED(rd.scalar(ps[i].zl));
        tr.common_scalar(ps[i].z);
        tr.common_scalar(ps[i].zn);
        if (ps[i].has_last) tr.common_scalar(ps[i].zl);
    }
    struct LEval { Fr z, zn, a, ap, s; };
    std::vector<LEval> le(s.lookups.size());
    for (auto& l : le) {                                                          // src/lookup.rs:108-171
        NEED(rscalar(l.z)); NEED(rscalar(l.zn)); NEED(rscalar(l.a)); NEED(rscalar(l.ap)); NEED(rscalar(l.s));
    }

    // x^n, l_0 / l_last / l_blind                                                 :513-591
    Fr xn = x;
    for (uint32_t i = 0; i < s.k; i++) xn = sqr(xn);
    Fr one = fr_one(), xn_m1 = xn - one, nfr = fr_from_u64(s.n);
    std::vector<Fr> l_evals;
    Fr wp = one;
    for (uint32_t i = 0; i < 2 + s.bf; i++) {
        l_evals.push_back(wp * xn_m1 * inv(nfr * (x - wp)));
        wp = wp * s.omega_inv;
    }
    Fr l_0 = l_evals[0], l_last = l_evals[1 + s.bf], l_blind = fr_zero();   // (before the reference's reverse())
    for (uint32_t i = 1; i <= s.bf; i++) l_blind = l_blind + l_evals[i];
    Fr one_minus = one - (l_last + l_blind);

    // expressions: gates, permutation (1..4), lookups (5 each)                    :593-645
    std::vector<Fr> ex;
    for (auto& g : s.gates) ex.push_back(eval_prog(g, s.consts, ae, fe, ie));
    if (s.n_chunks) {                                                             // src/permutation.rs:211-321
        ex.push_back(l_0 * (one - ps[0].z));
        Fr zl = ps.back().z;
        ex.push_back(l_last * (sqr(zl) - zl));
        for (uint32_t i = 1; i < s.n_chunks; i++) ex.push_back(l_0 * (ps[i].z - ps[i - 1].zl));
        Fr delta_pow = one;
        static const uint64_t DELTA_RAW[4] = {0x870e56bbe533e9a2ull, 0x5b5f898e5e963f25ull, 0x64ec26aad4c86e71ull,
                                              0x09226b6e22c6f0caull};
        Fr delta = fr_from_raw(DELTA_RAW);
        for (uint32_t ci = 0; ci < s.n_chunks; ci++) {
            Fr left = ps[ci].zn, right = ps[ci].z;
            for (uint32_t i = ci * s.chunk_len; i < std::min<size_t>((ci + 1) * s.chunk_len, s.perm.size()); i++) {
                const PermCol& pc = s.perm[i];
                Fr val = pc.type == COL_ADVICE ? ae[pc.qidx] : pc.type == COL_FIXED ? fe[pc.qidx] : ie[pc.qidx];
                left = left * (beta * se[i] + val + gamma);
                right = right * (beta * delta_pow * x + val + gamma);
                delta_pow = delta_pow * delta;
            }
            ex.push_back((left - right) * one_minus);
        }
    }
    for (size_t li = 0; li < s.lookups.size(); li++) {                            // src/lookup.rs:190-310
        const LEval& l = le[li];
        ex.push_back(l_0 * (one - l.z));
        ex.push_back(l_last * (sqr(l.z) - l.z));
        Fr ci = fr_zero(), ct = fr_zero();
        for (auto& g : s.lookups[li].inputs) ci = ci * theta + eval_prog(g, s.consts, ae, fe, ie);
        for (auto& g : s.lookups[li].tables) ct = ct * theta + eval_prog(g, s.consts, ae, fe, ie);
        Fr left = (l.a + beta) * (l.s + gamma) * l.zn, right = (ci + beta) * (ct + gamma) * l.z;
        ex.push_back((left - right) * one_minus);
        ex.push_back(l_0 * (l.a - l.s));
        ex.push_back((l.a - l.s) * (l.a - l.ap) * one_minus);
    }
    if (ex.empty()) H2A_FAIL(ctx, H2A_ERR_INVALID, "verify: the circuit has no constraint");
    Fr h_eval = ex[0];                                                            // src/vanishing.rs:145-175
    for (size_t i = 1; i < ex.size(); i++) h_eval = h_eval * y + ex[i];
    h_eval = h_eval * inv(xn_m1);

    // query list, in the reference's order                                        :654-715
    struct Q { int kind; size_t idx; int32_t rot; Fr eval; };  // kind 0 plain point in `pts`, 1 the composite H
    std::vector<PointA> pts;
    std::vector<Q> qs;
    auto add = [&](const PointA& p, int32_t rot, const Fr& ev) { pts.push_back(p); qs.push_back(Q{0, pts.size() - 1, rot, ev}); };
    for (size_t i = 0; i < s.iq.size(); i++) add(inst[s.iq[i].col], s.iq[i].rot, ie[i]);
    for (size_t i = 0; i < s.aq.size(); i++) add(adv[s.aq[i].col], s.aq[i].rot, ae[i]);
    for (uint32_t i = 0; i < s.n_chunks; i++) { add(pz[i], 0, ps[i].z); add(pz[i], 1, ps[i].zn); }   // src/permutation.rs:333-358
    for (int i = (int)s.n_chunks - 2; i >= 0; i--) add(pz[i], s.last_rot, ps[i].zl);
    for (size_t i = 0; i < s.lookups.size(); i++) {                               // src/lookup.rs:314-347
        add(lk_z[i], 0, le[i].z); add(lk_a[i], 0, le[i].a); add(lk_s[i], 0, le[i].s); add(lk_a[i], -1, le[i].ap); add(lk_z[i], 1, le[i].zn);
    }
    for (size_t i = 0; i < s.fq.size(); i++) add(affine_load(c->fixed_comms.data() + 64 * s.fq[i].col), s.fq[i].rot, fe[i]);
    for (size_t i = 0; i < s.perm.size(); i++) add(affine_load(c->sigma_comms.data() + 64 * i), 0, se[i]);
    qs.push_back(Q{1, 0, 0, h_eval});                                             // src/vanishing.rs:206-219
    add(random_comm, 0, random_eval);

    Fr v = tr.squeeze(), u = tr.squeeze();                                        // :718-719 (W_i are never absorbed)
    std::map<int32_t, int> rots;
    for (auto& q : qs) rots[q.rot] = 1;
    std::vector<uint8_t> ws(64 * rots.size());
    for (size_t i = 0; i < rots.size(); i++) {                                    // src/multiopen.rs:392
        PointA w;
        NEED(rd.point(w));
        affine_store(ws.data() + 64 * i, w);
    }
    if (rd.pos != len) H2A_FAIL(ctx, H2A_ERR_PROOF, "verify: %zu trailing bytes", len - rd.pos);
#undef NEED

    std::vector<int32_t> qrot(qs.size());
    std::vector<uint8_t> qev(32 * qs.size()), pbytes(64 * pts.size()), hbytes(64 * h.size());
    for (size_t i = 0; i < qs.size(); i++) { qrot[i] = qs[i].rot; fr_store(qev.data() + 32 * i, qs[i].eval); }
    for (size_t i = 0; i < pts.size(); i++) affine_store(pbytes.data() + 64 * i, pts[i]);
    for (size_t i = 0; i < h.size(); i++) affine_store(hbytes.data() + 64 * i, h[i]);
    h2a_glue::CommitmentEmitter emit = [&](h2a_glue::TermList& t, size_t q, const Fr& sc) {
        if (qs[q].kind == 0) { t.term(pbytes.data() + 64 * qs[q].idx, sc); return; }
        Fr p = sc;                                                                // H = sum_i (x^n)^i h_i, flattened
        for (size_t i = 0; i < h.size(); i++) { t.term(hbytes.data() + 64 * i, p); p = p * xn; }
    };
    uint8_t xb[32], ub[32], vb[32];
    fr_store(xb, x); fr_store(ub, u); fr_store(vb, v);
    if (!h2a_glue::expand_proof(tl, emit, qrot.data(), qev.data(), qs.size(), ws.data(), rots.size(), xb, ub, vb, s.omega,
                                s.omega_inv, g1))
        H2A_FAIL(ctx, H2A_ERR_PROOF, "verify: rotation sets and witness points disagree");
    return H2A_OK;
}

}  // namespace

extern "C" {

int h2a_circuit_create(h2a_ctx* ctx, const uint32_t* shape_words, size_t n_words, const uint8_t* constants, size_t n_constants,
                       h2a_circuit** out) {
    H2A_DEVICE(ctx);
    if (!ctx || !shape_words || !out || (!constants && n_constants)) return H2A_ERR_INVALID;
    h2a_circuit* c = new h2a_circuit();
    std::string err;
    if (!h2a_plonk::parse_shape(shape_words, n_words, constants, n_constants, c->shape, err)) {
        delete c;
        H2A_FAIL(ctx, H2A_ERR_INVALID, "%s", err.c_str());
    }
    *out = c;
    return H2A_OK;
}

int h2a_circuit_free(h2a_ctx* ctx, h2a_circuit* c) {
    H2A_DEVICE(ctx);
    if (!ctx || !c) return H2A_ERR_INVALID;
    if (c->prover) h2a_prover_state_free(ctx, c->prover);
    delete c;
    return H2A_OK;
}

int h2a_circuit_set_vk(h2a_ctx* ctx, h2a_circuit* c, const uint8_t* fixed_comms, const uint8_t* sigma_comms,
                       const uint8_t vk_hash[32]) {
    H2A_DEVICE(ctx);
    if (!ctx || !c || !vk_hash || (!fixed_comms && c->shape.n_fixed) || (!sigma_comms && !c->shape.perm.empty()))
        return H2A_ERR_INVALID;
    c->fixed_comms.assign(fixed_comms, fixed_comms + 64 * (size_t)c->shape.n_fixed);
    c->sigma_comms.assign(sigma_comms, sigma_comms + 64 * c->shape.perm.size());
    memcpy(c->vk_hash, vk_hash, 32);
    c->has_vk = true;
    return H2A_OK;
}

int h2a_verify_proof_batch(h2a_ctx* ctx, const h2a_circuit* c, size_t n_proofs, const uint8_t* inst_comms,
                           const uint8_t* const* proofs, const size_t* proof_lens, uint8_t* out_efwzw) {
    H2A_DEVICE(ctx);
    if (!ctx || !c || !out_efwzw || (n_proofs && (!proofs || !proof_lens)) || (c->shape.n_instance && !inst_comms))
        return H2A_ERR_INVALID;
    if (!c->has_vk) H2A_FAIL(ctx, H2A_ERR_INVALID, "verify: no verifying key set (h2a_circuit_set_vk / h2a_circuit_set_keys)");
    uint8_t g1[64];
    affine_store(g1, PointA{fq_one(), Fq{el_from_u64(2, MOD_Q)}});
    // transcript replay, point decompression (one square root per point) and expression evaluation are independent per
    // proof: spread them over the host cores, then evaluate all 4 * n_proofs sums in one launch
    std::vector<h2a_glue::TermList> lists(n_proofs);
    std::vector<ErrSink> errs(n_proofs);
    std::vector<int> rcs(n_proofs, H2A_OK);
    auto replay = [&](size_t first, size_t step) {
        for (size_t i = first; i < n_proofs; i += step)
            rcs[i] = collect_terms(&errs[i], c, inst_comms + 64 * (size_t)c->shape.n_instance * i, proofs[i], proof_lens[i], g1, lists[i]);
    };
    const size_t workers = n_proofs < 4 ? 1 : std::min<size_t>({n_proofs, (size_t)std::max(1u, std::thread::hardware_concurrency()), 32});
    if (workers <= 1) {
        replay(0, 1);
    } else {
        std::vector<std::thread> pool;
        for (size_t t = 1; t < workers; t++) pool.emplace_back(replay, t, workers);
        replay(0, workers);
        for (auto& th : pool) th.join();
    }
    h2a_glue::TermList tl;
    for (size_t i = 0; i < n_proofs; i++) {
        if (rcs[i] != H2A_OK) {
            ctx->err = errs[i].err;
            return rcs[i];
        }
        const uint32_t base = (uint32_t)(tl.bases.size() / 64);
        tl.bases.insert(tl.bases.end(), lists[i].bases.begin(), lists[i].bases.end());
        tl.scalars.insert(tl.scalars.end(), lists[i].scalars.begin(), lists[i].scalars.end());
        for (size_t q = 1; q < lists[i].offsets.size(); q++) tl.offsets.push_back(base + lists[i].offsets[q]);
    }
    return h2a_small_msm(ctx, tl.bases.data(), tl.scalars.data(), tl.offsets.data(), 4 * n_proofs, out_efwzw);
}

int h2a_verify_proof(h2a_ctx* ctx, const h2a_circuit* c, const uint8_t* inst_comms, const uint8_t* proof, size_t proof_len,
                     uint8_t out_efwzw[256]) {
    H2A_DEVICE(ctx);
    const uint8_t* proofs[1] = {proof};
    size_t lens[1] = {proof_len};
    return h2a_verify_proof_batch(ctx, c, 1, inst_comms, proofs, lens, out_efwzw);
}

}  // extern "C"
