// glue.cu — the verifier-side glue of include/h2agg.h: the Blake2b transcript, the expansion of the
// GWC multi-open accumulation into flat (scalar, base) term lists, and the batched small-MSM kernel
// that evaluates all of those sums in one launch.
//
// Reference semantics being reproduced (native values only; the constraint system is out of scope):
//   MultiopenChip::calc_witness   src/multiopen.rs:271-509   (w, zw, f, e)
//   construct_intermediate_sets   src/multiopen.rs:19-45     (BTreeMap order of rotations)
//   VanishingChip::verify H fold  src/vanishing.rs:177-188
//   TranscriptChip                src/transcript.rs:66-129   (Blake2bWrite + Challenge255)
#include <algorithm>
#include <cstring>
#include <map>
#include <vector>

#include "ctx.hpp"
#include "curve.cuh"
#include "host_bn254.hpp"
#include "host_glue.hpp"

namespace dev {
using namespace h2a;

__device__ __noinline__ void xyzz_add_fn(XYZZ& a, const XYZZ& b) { a.add(b); }

// One block per output sum; each thread multiplies its terms with signed 4-bit windows over the canonical scalar
// (8-entry table, 65 windows of 4 doublings + at most one addition: the chain of ~256 doublings is the latency floor
// of a variable-base product), the block folds the partial sums in shared memory, thread 0 normalises to affine.
constexpr int SMALL_THREADS = 64;
__global__ void __launch_bounds__(SMALL_THREADS) small_msm_kernel(const uint8_t* __restrict__ bases,
                                                                  const uint8_t* __restrict__ scalars,
                                                                  const uint32_t* __restrict__ sum_offsets,
                                                                  uint8_t* __restrict__ out_affine) {
    __shared__ __align__(16) uint8_t sh[SMALL_THREADS * 128];
    const uint32_t lo = sum_offsets[blockIdx.x], hi = sum_offsets[blockIdx.x + 1];
    XYZZ acc = XYZZ::identity();
    for (uint32_t i = lo + threadIdx.x; i < hi; i += SMALL_THREADS) {
        Affine p = Affine::load(bases + 64ull * i);
        if (p.is_identity()) continue;
        Fr s = Fr::load(scalars + 32ull * i).from_mont();
        // signed digits d_w in [-8, 8), s = sum_w d_w 16^w; the table and the digits live in local memory by design
        int8_t dig[65];
        int carry = 0, top = -1;
        for (int w = 0; w < 64; w++) {
            int d = (int)((s.l[w >> 3] >> (4 * (w & 7))) & 15u) + carry;
            carry = d >= 8;
            d -= carry << 4;
            dig[w] = (int8_t)d;
            if (d) top = w;
        }
        dig[64] = (int8_t)carry;
        if (carry) top = 64;
        XYZZ tab[8];   // tab[k] = (k + 1) P
        tab[0] = XYZZ::from_affine(p);
        for (int k = 1; k < 8; k++) {
            tab[k] = tab[k - 1];
            xyzz_add_fn(tab[k], tab[0]);
        }
        XYZZ m = XYZZ::identity();
        for (int w = top; w >= 0; w--) {
            if (w != top)
                for (int k = 0; k < 4; k++) m = m.dbl();
            const int d = dig[w];
            if (d) {
                XYZZ t = tab[(d > 0 ? d : -d) - 1];
                if (d < 0) t.y = t.y.neg();
                xyzz_add_fn(m, t);
            }
        }
        xyzz_add_fn(acc, m);
    }
    acc.store(sh + 128 * threadIdx.x);
    __syncthreads();
    for (int stride = SMALL_THREADS / 2; stride > 0; stride >>= 1) {
        if ((int)threadIdx.x < stride) {
            XYZZ a = XYZZ::load(sh + 128 * threadIdx.x), b = XYZZ::load(sh + 128 * (threadIdx.x + stride));
            xyzz_add_fn(a, b);
            a.store(sh + 128 * threadIdx.x);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        XYZZ r = XYZZ::load(sh);
        Affine o;
        if (r.is_identity()) {
            o.x = Fq::zero();
            o.y = Fq::zero();
        } else {
            Fq zi = r.zzz.inv();
            Fq zzi = (zi * r.zz).sqr();
            o.x = r.x * zzi;
            o.y = r.y * zi;
        }
        o.store(out_affine + 64ull * blockIdx.x);
    }
}

}  // namespace dev
using dev::SMALL_THREADS;
using dev::small_msm_kernel;

// n_sums independent sums: sum s covers terms [sum_offsets[s], sum_offsets[s+1]).  Host pointers.
int h2a_small_msm(h2a_ctx* ctx, const uint8_t* bases, const uint8_t* scalars, const uint32_t* sum_offsets, size_t n_sums,
                  uint8_t* out_affine) {
    H2A_DEVICE(ctx);
    if (n_sums == 0) return H2A_OK;
    const size_t n_terms = sum_offsets[n_sums];
    const size_t off_bytes = (n_sums + 1) * 4;
    const size_t total = n_terms * 96 + off_bytes + n_sums * 64 + 256;
    H2A_TRY(h2a_reserve(ctx, ctx->heavy, total));
    H2A_TRY(h2a_reserve_pinned(ctx, n_sums * 64));
    uint8_t* d = (uint8_t*)ctx->heavy.p;
    uint8_t* d_bases = d;
    uint8_t* d_scal = d_bases + n_terms * 64;
    uint8_t* d_offs = d_scal + n_terms * 32;
    d_offs += (16 - ((uintptr_t)d_offs & 15)) & 15;
    uint8_t* d_out = d_offs + ((off_bytes + 15) & ~(size_t)15);
    if (n_terms) {
        H2A_CUDA(ctx, cudaMemcpyAsync(d_bases, bases, n_terms * 64, cudaMemcpyHostToDevice, ctx->stream));
        H2A_CUDA(ctx, cudaMemcpyAsync(d_scal, scalars, n_terms * 32, cudaMemcpyHostToDevice, ctx->stream));
    }
    H2A_CUDA(ctx, cudaMemcpyAsync(d_offs, sum_offsets, off_bytes, cudaMemcpyHostToDevice, ctx->stream));
    small_msm_kernel<<<(unsigned)n_sums, SMALL_THREADS, 0, ctx->stream>>>(d_bases, d_scal, (const uint32_t*)d_offs, d_out);
    H2A_LAUNCH_CHECK(ctx);
    H2A_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, d_out, n_sums * 64, cudaMemcpyDeviceToHost, ctx->stream));
    H2A_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(out_affine, ctx->pinned, n_sums * 64);
    return H2A_OK;
}

struct h2a_transcript {
    h2a_glue::Transcript t;
};

namespace h2a_glue {

// Appends the four sums (e, f, w, zw) of one proof.  Returns false if n_ws != number of rotation sets.
bool expand_proof(TermList& tl, const CommitmentEmitter& emit, const int32_t* rotations, const uint8_t* evals, size_t nq,
                  const uint8_t* ws, size_t n_ws, const uint8_t* x_, const uint8_t* u_, const uint8_t* v_,
                  const h2a_host::Fr& omega, const h2a_host::Fr& omega_inv, const uint8_t* g1) {
    using namespace h2a_host;
    std::map<int32_t, std::vector<size_t>> sets;  // ascending rotation, insertion order inside
    for (size_t i = 0; i < nq; i++) sets[rotations[i]].push_back(i);
    const size_t S = sets.size();
    if (S != n_ws || S == 0) return false;
    Fr x = fr_load(x_), u = fr_load(u_), v = fr_load(v_);
    std::vector<Fr> upow(S);  // upow[i] = u^(S-1-i)
    upow[S - 1] = fr_one();
    for (size_t i = S - 1; i-- > 0;) upow[i] = upow[i + 1] * u;

    struct FTerm { size_t q; Fr s; };
    std::vector<FTerm> fterms;
    std::vector<Fr> zs(S);
    Fr e_scalar = fr_zero();
    size_t si = 0;
    for (auto& kv : sets) {
        int32_t rot = kv.first;
        Fr om = rot >= 0 ? pow_u64(omega, (uint64_t)rot) : pow_u64(omega_inv, (uint64_t)(-(int64_t)rot));
        zs[si] = om * x;
        const std::vector<size_t>& qs = kv.second;
        Fr vp = fr_one();  // v^(m-1-j), built from the last query backwards
        Fr ev = fr_zero();
        std::vector<Fr> coef(qs.size());
        for (size_t j = qs.size(); j-- > 0;) {
            coef[j] = vp * upow[si];
            ev = ev + vp * fr_load(evals + 32 * qs[j]);
            vp = vp * v;
        }
        for (size_t j = 0; j < qs.size(); j++) fterms.push_back(FTerm{qs[j], coef[j]});
        e_scalar = e_scalar + upow[si] * ev;
        si++;
    }
    tl.term(g1, neg(e_scalar));                                                       // E
    tl.close_sum();
    for (auto& t : fterms) emit(tl, t.q, t.s);                                        // F
    tl.close_sum();
    for (size_t i = 0; i < S; i++) tl.term(ws + 64 * i, upow[i]);                     // W
    tl.close_sum();
    for (size_t i = 0; i < S; i++) tl.term(ws + 64 * i, upow[i] * zs[i]);             // ZW
    tl.close_sum();
    return true;
}


}  // namespace h2a_glue
using h2a_glue::TermList;
using h2a_glue::expand_proof;

extern "C" {

// rand_xorshift 0.3 XorShiftRng::from_seed + 64 bytes of fill_bytes -> Fr::from_bytes_wide: the draw that
// `Setup::<Bn256>::new(k, XorShiftRng::from_seed(seed))` makes for the KZG secret (examples/simple-example.rs:584-589).
// The Fr::random convention is upstream-inferred; the stream itself is the published algorithm.
int h2a_xorshift_scalar(const uint8_t seed[16], uint8_t out_scalar[32]) {
    if (!seed || !out_scalar) return H2A_ERR_INVALID;
    uint32_t st[4];
    memcpy(st, seed, 16);
    if ((st[0] | st[1] | st[2] | st[3]) == 0) st[0] = st[1] = st[2] = st[3] = 0x0BAD5EEDu;
    uint8_t wide[64];
    for (int i = 0; i < 16; i++) {
        uint32_t t = st[0] ^ (st[0] << 11);
        st[0] = st[1]; st[1] = st[2]; st[2] = st[3];
        st[3] = st[3] ^ (st[3] >> 19) ^ (t ^ (t >> 8));
        memcpy(wide + 4 * i, &st[3], 4);
    }
    h2a_host::fr_store(out_scalar, h2a_glue::fr_from_wide(wide));
    return H2A_OK;
}

// `vk.pinned()` hashed into the transcript (src/verifier.rs:341-358): Blake2b-512, personal "Halo2-Verify-Key", over
// len_le64 || bytes of the Debug string, then Fr::from_bytes_wide.
int h2a_vk_hash(const uint8_t* pinned_debug, size_t len, uint8_t out_scalar[32]) {
    if (!out_scalar || (!pinned_debug && len)) return H2A_ERR_INVALID;
    h2a_glue::Blake2bState st(64, "Halo2-Verify-Key");
    uint8_t le[8];
    const uint64_t l64 = (uint64_t)len;
    memcpy(le, &l64, 8);
    st.absorb(le, 8);
    if (len) st.absorb(pinned_debug, len);
    uint8_t wide[64];
    st.digest(wide);
    h2a_host::fr_store(out_scalar, h2a_glue::fr_from_wide(wide));
    return H2A_OK;
}

h2a_transcript* h2a_transcript_new(void) { return new h2a_transcript(); }
void h2a_transcript_free(h2a_transcript* t) { delete t; }
int h2a_transcript_common_point(h2a_transcript* t, const uint8_t point_affine[64]) {
    if (!t || !point_affine) return H2A_ERR_INVALID;
    return t->t.common_point(h2a_host::affine_load(point_affine)) ? H2A_OK : H2A_ERR_INVALID;
}
int h2a_transcript_common_scalar(h2a_transcript* t, const uint8_t scalar[32]) {
    if (!t || !scalar) return H2A_ERR_INVALID;
    t->t.common_scalar(h2a_host::fr_load(scalar));
    return H2A_OK;
}
int h2a_transcript_squeeze_challenge(h2a_transcript* t, uint8_t out_scalar[32]) {
    if (!t || !out_scalar) return H2A_ERR_INVALID;
    h2a_host::fr_store(out_scalar, t->t.squeeze());
    return H2A_OK;
}

int h2a_verify_accumulate_batch(h2a_ctx* ctx, size_t n_proofs, const uint8_t* commitments, const int32_t* rotations,
                                const uint8_t* evals, const size_t* q_off, const uint8_t* ws, const size_t* w_off,
                                const uint8_t* xuv, const uint8_t omega[32], const uint8_t g1[64], uint8_t* out_efwzw) {
    H2A_DEVICE(ctx);
    if (!ctx || !commitments || !rotations || !evals || !q_off || !ws || !w_off || !xuv || !omega || !g1 || !out_efwzw)
        return H2A_ERR_INVALID;
    using namespace h2a_host;
    Fr om = fr_load(omega), om_inv = inv(om);
    TermList tl;
    for (size_t p = 0; p < n_proofs; p++) {
        size_t q0 = q_off[p], q1 = q_off[p + 1], w0 = w_off[p], w1 = w_off[p + 1];
        if (q1 < q0 || w1 < w0) H2A_FAIL(ctx, H2A_ERR_INVALID, "verify_accumulate: offsets of proof %zu not monotone", p);
        const uint8_t* cbase = commitments + 64 * q0;
        h2a_glue::CommitmentEmitter emit = [cbase](TermList& t, size_t q, const Fr& sc) { t.term(cbase + 64 * q, sc); };
        if (!expand_proof(tl, emit, rotations + q0, evals + 32 * q0, q1 - q0, ws + 64 * w0, w1 - w0,
                          xuv + 96 * p, xuv + 96 * p + 32, xuv + 96 * p + 64, om, om_inv, g1))
            H2A_FAIL(ctx, H2A_ERR_INVALID, "verify_accumulate: proof %zu has %zu W points for a different number of rotation sets",
                     p, w1 - w0);
    }
    return h2a_small_msm(ctx, tl.bases.data(), tl.scalars.data(), tl.offsets.data(), 4 * n_proofs, out_efwzw);
}

int h2a_verify_accumulate(h2a_ctx* ctx, const uint8_t* commitments, const int32_t* rotations, const uint8_t* evals, size_t nq,
                          const uint8_t* ws, size_t n_ws, const uint8_t x[32], const uint8_t u[32], const uint8_t v[32],
                          const uint8_t omega[32], const uint8_t g1[64], uint8_t out_efwzw[256]) {
    H2A_DEVICE(ctx);
    if (!x || !u || !v) return H2A_ERR_INVALID;
    size_t q_off[2] = {0, nq}, w_off[2] = {0, n_ws};
    uint8_t xuv[96];
    memcpy(xuv, x, 32);
    memcpy(xuv + 32, u, 32);
    memcpy(xuv + 64, v, 32);
    return h2a_verify_accumulate_batch(ctx, 1, commitments, rotations, evals, q_off, ws, w_off, xuv, omega, g1, out_efwzw);
}

int h2a_fold_h(h2a_ctx* ctx, const uint8_t* h_pieces, size_t m, const uint8_t xn[32], uint8_t out_affine[64]) {
    H2A_DEVICE(ctx);
    if (!ctx || !h_pieces || !xn || !out_affine || m == 0) return H2A_ERR_INVALID;
    using namespace h2a_host;
    TermList tl;
    Fr p = fr_one(), x = fr_load(xn);
    for (size_t i = 0; i < m; i++) {
        tl.term(h_pieces + 64 * i, p);
        p = p * x;
    }
    tl.close_sum();
    return h2a_small_msm(ctx, tl.bases.data(), tl.scalars.data(), tl.offsets.data(), 1, out_affine);
}

}  // extern "C"
