// host_glue.hpp — host-side pieces shared by the verifier glue and the prover: the Blake2b transcript
// (src/transcript.rs:66-129), Fr::from_bytes_wide, and the expansion of the GWC accumulation
// (src/multiopen.rs:19-45,271-509) into flat (scalar, base) term lists for the small-MSM kernel.
#pragma once
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <vector>

#include "host_bn254.hpp"

namespace h2a_glue {

class Blake2bState {
  public:
    Blake2bState(unsigned out_len, const char personal[16]) : out_len_(out_len) {
        static const uint64_t iv[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull,
                                       0xa54ff53a5f1d36f1ull, 0x510e527fade682d1ull, 0x9b05688c2b3e6c1full,
                                       0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
        memcpy(iv_, iv, sizeof iv);
        memcpy(h_, iv, sizeof iv);
        h_[0] ^= 0x01010000ull ^ out_len;  // fanout = depth = 1, no key
        uint64_t p0, p1;
        memcpy(&p0, personal, 8);
        memcpy(&p1, personal + 8, 8);
        h_[6] ^= p0;
        h_[7] ^= p1;
    }
    void absorb(const uint8_t* data, size_t len) {
        while (len) {
            if (fill_ == 128) {  // a full block is only compressed once more input follows it
                counter_ += 128;
                round_block(false);
                fill_ = 0;
            }
            const size_t take = len < 128 - fill_ ? len : 128 - fill_;
            memcpy(block_ + fill_, data, take);
            fill_ += take;
            data += take;
            len -= take;
        }
    }
    void digest(uint8_t* out) const {  // does not disturb the running state
        Blake2bState c = *this;
        c.counter_ += c.fill_;
        memset(c.block_ + c.fill_, 0, 128 - c.fill_);
        c.round_block(true);
        memcpy(out, c.h_, out_len_);
    }

  private:
    static uint64_t ror(uint64_t v, int r) { return (v >> r) | (v << (64 - r)); }
    void round_block(bool final) {
        static const uint8_t sigma[10][16] = {
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
            {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
            {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
            {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
            {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};
        uint64_t m[16], v[16];
        memcpy(m, block_, 128);
        for (int i = 0; i < 8; i++) {
            v[i] = h_[i];
            v[i + 8] = iv_[i];
        }
        v[12] ^= counter_;  // 64-bit byte counter is ample for a transcript
        if (final) v[14] = ~v[14];
        for (int r = 0; r < 12; r++) {
            const uint8_t* s = sigma[r % 10];
            for (int g = 0; g < 8; g++) {
                int a, b, c, d;
                if (g < 4) { a = g; b = 4 + g; c = 8 + g; d = 12 + g; }
                else { a = g - 4; b = 4 + (g - 3) % 4; c = 8 + (g - 2) % 4; d = 12 + (g - 1) % 4; }
                v[a] += v[b] + m[s[2 * g]];
                v[d] = ror(v[d] ^ v[a], 32);
                v[c] += v[d];
                v[b] = ror(v[b] ^ v[c], 24);
                v[a] += v[b] + m[s[2 * g + 1]];
                v[d] = ror(v[d] ^ v[a], 16);
                v[c] += v[d];
                v[b] = ror(v[b] ^ v[c], 63);
            }
        }
        for (int i = 0; i < 8; i++) h_[i] ^= v[i] ^ v[i + 8];
    }
    uint64_t h_[8], iv_[8];
    uint8_t block_[128];
    size_t fill_ = 0;
    uint64_t counter_ = 0;
    unsigned out_len_;
};

inline h2a_host::Fr fr_from_wide(const uint8_t b[64]) {  // 512-bit LE integer mod r -> Montgomery
    using namespace h2a_host;
    El lo, hi, r2;
    memcpy(lo.v, b, 32);
    memcpy(hi.v, b + 32, 32);
    memcpy(r2.v, MOD_R.r2, 32);
    El r3 = el_mul(r2, r2, MOD_R);
    // el_mul tolerates one unreduced (< 2^256) operand: the result stays below 2r before the final subtraction
    return Fr{el_add(el_mul(lo, r2, MOD_R), el_mul(hi, r3, MOD_R), MOD_R)};
}

// Blake2bWrite / Blake2bRead with Challenge255: prefix 1 = point, 2 = scalar, 0 = squeeze.
struct Transcript {
    Blake2bState st{64, "Halo2-Transcript"};
    bool common_point(const h2a_host::PointA& p) {
        using namespace h2a_host;
        if (is_identity(p)) return false;  // halo2 cannot absorb the identity's coordinates
        uint8_t buf[65];
        buf[0] = 1;
        uint64_t raw[4];
        fq_to_raw(p.x, raw);
        memcpy(buf + 1, raw, 32);
        fq_to_raw(p.y, raw);
        memcpy(buf + 33, raw, 32);
        st.absorb(buf, 65);
        return true;
    }
    void common_scalar(const h2a_host::Fr& s) {
        uint8_t buf[33];
        buf[0] = 2;
        uint64_t raw[4];
        h2a_host::fr_to_raw(s, raw);
        memcpy(buf + 1, raw, 32);
        st.absorb(buf, 33);
    }
    h2a_host::Fr squeeze() {
        uint8_t zero = 0;
        st.absorb(&zero, 1);
        uint8_t wide[64];
        st.digest(wide);
        return fr_from_wide(wide);
    }
};

struct TermList {
    std::vector<uint8_t> bases, scalars;
    std::vector<uint32_t> offsets{0};
    void term(const uint8_t* base64, const h2a_host::Fr& s) {
        bases.insert(bases.end(), base64, base64 + 64);
        uint8_t b[32];
        h2a_host::fr_store(b, s);
        scalars.insert(scalars.end(), b, b + 32);
    }
    void close_sum() { offsets.push_back((uint32_t)(bases.size() / 64)); }
};

// emit(tl, q, s) appends the term(s) s * C_q of query q to the list: one term for a plain commitment,
// several for a composite one (H = sum_i (x^n)^i h_i is flattened into its pieces, src/vanishing.rs:177-188).
typedef std::function<void(TermList&, size_t, const h2a_host::Fr&)> CommitmentEmitter;

// Appends the four sums (e, f, w, zw) of one proof.  Returns false if n_ws != number of rotation sets.
bool expand_proof(TermList& tl, const CommitmentEmitter& emit, const int32_t* rotations, const uint8_t* evals, size_t nq,
                  const uint8_t* ws, size_t n_ws, const uint8_t* x_, const uint8_t* u_, const uint8_t* v_,
                  const h2a_host::Fr& omega, const h2a_host::Fr& omega_inv, const uint8_t* g1);

}  // namespace h2a_glue
