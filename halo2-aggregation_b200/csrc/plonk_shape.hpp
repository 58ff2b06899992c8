// plonk_shape.hpp — the description of a circuit that `VerifierChip::_verify_proof` receives from the
// verifying key (src/verifier.rs:286-311: log_n, blinding_factors, column counts, lookups, permutation columns
// and chunk length, gates, query lists), as a flat word stream that crosses the C ABI (h2a_circuit_create).
//
// word stream: MAGIC, k, blinding_factors, degree, n_instance, n_advice, n_fixed,
//              n_aq, (col, rot)*, n_fq, (col, rot)*, n_iq, (col, rot)*,
//              n_gates, (len, (op, arg)*)*, n_constants,
//              n_lookups, (n_inputs, (len, (op, arg)*)*, n_tables, (len, (op, arg)*)*)*,
//              n_perm, (type, col, query_index)*
// ops (the `Expression` tree of src/verifier.rs:58-151 in postfix order):
//   0 Constant(c[arg]) 1 Advice(query arg) 2 Fixed(query arg) 3 Instance(query arg) 4 Negated 5 Sum 6 Product 7 Scaled(c[arg])
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "host_bn254.hpp"

namespace h2a_plonk {

constexpr uint32_t SHAPE_MAGIC = 0x48324153u;
enum Op : uint32_t { OP_CONST = 0, OP_ADVICE, OP_FIXED, OP_INSTANCE, OP_NEG, OP_ADD, OP_MUL, OP_SCALE };
enum ColType : uint32_t { COL_ADVICE = 0, COL_FIXED = 1, COL_INSTANCE = 2 };

struct Query { uint32_t col; int32_t rot; };
struct Prog { std::vector<uint32_t> code; };  // (op, arg) pairs
struct Lookup { std::vector<Prog> inputs, tables; };
struct PermCol { uint32_t type, col, qidx; };

struct Shape {
    uint32_t k = 0, bf = 0, degree = 0, n_instance = 0, n_advice = 0, n_fixed = 0;
    std::vector<Query> aq, fq, iq;
    std::vector<Prog> gates;
    std::vector<h2a_host::Fr> consts;
    std::vector<Lookup> lookups;
    std::vector<PermCol> perm;
    // derived
    uint32_t n = 0, chunk_len = 0, qdeg = 0, ext_k = 0, usable = 0, n_chunks = 0;
    int32_t last_rot = 0;
    h2a_host::Fr omega, omega_inv;
};

inline bool parse_shape(const uint32_t* w, size_t nw, const uint8_t* consts, size_t n_consts, Shape& s, std::string& err) {
    size_t p = 0;
    auto need = [&](size_t c) { return p + c <= nw; };
    auto bad = [&](const char* m) { err = m; return false; };
    if (!need(7) || w[0] != SHAPE_MAGIC) return bad("shape: bad magic or truncated header");
    s.k = w[1]; s.bf = w[2]; s.degree = w[3]; s.n_instance = w[4]; s.n_advice = w[5]; s.n_fixed = w[6];
    p = 7;
    if (s.k < 2 || s.k > 26 || s.degree < 3 || s.degree > 9) return bad("shape: k or degree out of range");
    auto queries = [&](std::vector<Query>& q, uint32_t ncols) {
        if (!need(1)) return false;
        uint32_t c = w[p++];
        if (!need(2 * (size_t)c)) return false;
        for (uint32_t i = 0; i < c; i++) {
            Query x{w[p], (int32_t)w[p + 1]};
            p += 2;
            if (x.col >= ncols) return false;
            q.push_back(x);
        }
        return true;
    };
    if (!queries(s.aq, s.n_advice) || !queries(s.fq, s.n_fixed) || !queries(s.iq, s.n_instance)) return bad("shape: bad query list");
    auto prog = [&](Prog& g) {
        if (!need(1)) return false;
        uint32_t len = w[p++];
        if (!need(2 * (size_t)len)) return false;
        int depth = 0;
        for (uint32_t i = 0; i < len; i++) {
            uint32_t op = w[p], arg = w[p + 1];
            p += 2;
            switch (op) {
                case OP_CONST: if (arg >= n_consts) return false; depth++; break;
                case OP_ADVICE: if (arg >= s.aq.size()) return false; depth++; break;
                case OP_FIXED: if (arg >= s.fq.size()) return false; depth++; break;
                case OP_INSTANCE: if (arg >= s.iq.size()) return false; depth++; break;
                case OP_NEG: if (depth < 1) return false; break;
                case OP_ADD: case OP_MUL: if (depth < 2) return false; depth--; break;
                case OP_SCALE: if (depth < 1 || arg >= n_consts) return false; break;
                default: return false;
            }
            if (depth > 16) return false;
            g.code.push_back(op);
            g.code.push_back(arg);
        }
        return depth == 1;
    };
    if (!need(1)) return bad("shape: truncated");
    uint32_t ng = w[p++];
    for (uint32_t i = 0; i < ng; i++) {
        Prog g;
        if (!prog(g)) return bad("shape: bad gate program");
        s.gates.push_back(g);
    }
    if (!need(1) || w[p++] != n_consts) return bad("shape: constant count mismatch");
    for (size_t i = 0; i < n_consts; i++) s.consts.push_back(h2a_host::fr_load(consts + 32 * i));
    if (!need(1)) return bad("shape: truncated");
    uint32_t nl = w[p++];
    for (uint32_t i = 0; i < nl; i++) {
        Lookup lk;
        for (int side = 0; side < 2; side++) {
            if (!need(1)) return bad("shape: truncated lookup");
            uint32_t c = w[p++];
            for (uint32_t j = 0; j < c; j++) {
                Prog g;
                if (!prog(g)) return bad("shape: bad lookup program");
                (side ? lk.tables : lk.inputs).push_back(g);
            }
        }
        s.lookups.push_back(lk);
    }
    if (!need(1)) return bad("shape: truncated");
    uint32_t np = w[p++];
    if (!need(3 * (size_t)np)) return bad("shape: truncated permutation");
    for (uint32_t i = 0; i < np; i++) {
        PermCol c{w[p], w[p + 1], w[p + 2]};
        p += 3;
        size_t nq = c.type == COL_ADVICE ? s.aq.size() : c.type == COL_FIXED ? s.fq.size() : s.iq.size();
        if (c.type > 2 || c.qidx >= nq) return bad("shape: bad permutation column");
        // the query must be the one halo2's get_any_query_index(column, Rotation::cur()) returns: this column at rotation 0
        const Query& q = (c.type == COL_ADVICE ? s.aq : c.type == COL_FIXED ? s.fq : s.iq)[c.qidx];
        if (q.col != c.col || q.rot != 0) return bad("shape: permutation column does not name its own rotation-0 query");
        s.perm.push_back(c);
    }
    if (p != nw) return bad("shape: trailing words");
    s.n = 1u << s.k;
    if (s.bf + 2 >= s.n) return bad("shape: too many blinding factors");
    s.chunk_len = s.degree - 2;                    // src/verifier.rs:236
    s.qdeg = s.degree - 1;                         // get_quotient_poly_degree, src/verifier.rs:431
    s.ext_k = s.k;
    while ((1ull << s.ext_k) < (uint64_t)s.n * s.qdeg) s.ext_k++;
    s.usable = s.n - (s.bf + 1);
    s.n_chunks = (uint32_t)((s.perm.size() + s.chunk_len - 1) / s.chunk_len);
    s.last_rot = -(int32_t)(s.bf + 1);             // src/permutation.rs:335
    s.omega = h2a_host::fr_root_of_unity((int)s.k);
    s.omega_inv = h2a_host::inv(s.omega);
    return true;
}

inline h2a_host::Fr eval_prog(const Prog& g, const std::vector<h2a_host::Fr>& consts, const std::vector<h2a_host::Fr>& adv,
                              const std::vector<h2a_host::Fr>& fix, const std::vector<h2a_host::Fr>& inst) {
    using namespace h2a_host;
    Fr st[17];
    int sp = 0;
    for (size_t i = 0; i < g.code.size(); i += 2) {
        uint32_t op = g.code[i], arg = g.code[i + 1];
        switch (op) {
            case OP_CONST: st[sp++] = consts[arg]; break;
            case OP_ADVICE: st[sp++] = adv[arg]; break;
            case OP_FIXED: st[sp++] = fix[arg]; break;
            case OP_INSTANCE: st[sp++] = inst[arg]; break;
            case OP_NEG: st[sp - 1] = neg(st[sp - 1]); break;
            case OP_ADD: st[sp - 2] = st[sp - 2] + st[sp - 1]; sp--; break;
            case OP_MUL: st[sp - 2] = st[sp - 2] * st[sp - 1]; sp--; break;
            default: st[sp - 1] = st[sp - 1] * consts[arg]; break;  // OP_SCALE
        }
    }
    return st[0];
}

}  // namespace h2a_plonk
