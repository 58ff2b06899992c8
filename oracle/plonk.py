"""Oracle restatement of the PLONK + KZG (GWC multi-open) prover and verifier whose wire format the
reference's in-circuit verifier re-reads.  TEST INFRASTRUCTURE ONLY (see oracle/bn254.hpp header).

PARITY UNPINNED: `create_proof` / `verify_proof` live in the un-vendored `halo2` dependency
(Cargo.toml:12).  What IS pinned by the reference, and followed line by line here:
  * proof wire order and transcript absorption order      src/verifier.rs:341-510, src/lookup.rs:49-171,
                                                           src/permutation.rs:48-179, src/vanishing.rs:54-134
  * challenge order theta, beta, gamma, y, x, v, u        src/verifier.rs:378,390,393,423,436,718,719
  * l_0 / l_last / l_blind                                 src/verifier.rs:513-591
  * gate / permutation / lookup expressions and their order src/verifier.rs:593-645, src/permutation.rs:190-324,
                                                           src/lookup.rs:173-311
  * expected h(x) and the H fold                           src/vanishing.rs:145-188
  * query order and multi-open accumulation                src/verifier.rs:654-715, src/multiopen.rs:19-45,271-509
The prover side (what the dependency's `create_proof` must do for that verifier to accept) follows the
published halo2 v0.1.0-beta algorithm; randomness is injected through a
seeded counter-based stream so proofs are reproducible byte for byte.

Orchestration is plain Python integers; MSMs go through the C++ oracle (`orc.msm`).
"""
from . import pymodel as pm

R = pm.R
ADVICE, FIXED, INSTANCE = 0, 1, 2
OP_CONST, OP_ADVICE, OP_FIXED, OP_INSTANCE, OP_NEG, OP_ADD, OP_MUL, OP_SCALE = range(8)
DELTA = pow(7, 1 << 28, R)  # Fr::DELTA (src/permutation.rs:259)


class Shape:
    """What `_verify_proof` receives from the verifying key (src/verifier.rs:286-311)."""

    def __init__(self, k, blinding_factors, degree, num_instance, num_advice, num_fixed, advice_queries, fixed_queries,
                 instance_queries, gates, constants, lookups, perm_columns, coset_shift=7):
        self.k, self.n = k, 1 << k
        self.bf = blinding_factors
        self.degree = degree                      # cs.degree()
        self.chunk_len = degree - 2               # src/verifier.rs:236
        self.quotient_poly_degree = degree - 1    # EvaluationDomain::get_quotient_poly_degree (:431)
        ext = k
        while (1 << ext) < self.n * (degree - 1):
            ext += 1
        self.ext_k = ext
        self.num_instance, self.num_advice, self.num_fixed = num_instance, num_advice, num_fixed
        self.advice_queries, self.fixed_queries, self.instance_queries = advice_queries, fixed_queries, instance_queries
        self.gates = gates                        # list of RPN programs [(op, arg), ...]
        self.constants = constants                # canonical integers referenced by OP_CONST / OP_SCALE
        self.lookups = lookups                    # list of (input_programs, table_programs)
        self.perm_columns = perm_columns          # [(column type, column index, query index at rotation 0)]
        self.coset_shift = coset_shift
        self.omega = pm.omega_for(k)
        self.usable = self.n - (self.bf + 1)      # index of the l_last row


# ------------------------------------------------------------------ small helpers
def inv(a):
    return pow(a, -1, R)


def batch_inv(vals):
    return [inv(v) if v else 0 for v in vals]


def ntt(a, omega):
    """iterative radix-2; natural order in and out (value of best_fft)."""
    n = len(a)
    k = n.bit_length() - 1
    a = [a[int(format(i, "0%db" % k)[::-1], 2)] if k else a[i] for i in range(n)]
    m = 1
    while m < n:
        wm = pow(omega, n // (2 * m), R)
        for s in range(0, n, 2 * m):
            w = 1
            for j in range(m):
                t = a[s + j + m] * w % R
                u = a[s + j]
                a[s + j] = (u + t) % R
                a[s + j + m] = (u - t) % R
                w = w * wm % R
        m *= 2
    return a


def intt(a, omega):
    n = len(a)
    ninv = inv(n)
    return [v * ninv % R for v in ntt(a, inv(omega))]


def poly_eval(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % R
    return acc


def eval_rpn(prog, consts, adv, fix, inst):
    st = []
    for op, arg in prog:
        if op == OP_CONST:
            st.append(consts[arg] % R)
        elif op == OP_ADVICE:
            st.append(adv[arg])
        elif op == OP_FIXED:
            st.append(fix[arg])
        elif op == OP_INSTANCE:
            st.append(inst[arg])
        elif op == OP_NEG:
            st.append((-st.pop()) % R)
        elif op == OP_ADD:
            b = st.pop(); a = st.pop(); st.append((a + b) % R)
        elif op == OP_MUL:
            b = st.pop(); a = st.pop(); st.append(a * b % R)
        elif op == OP_SCALE:
            st.append(st.pop() * consts[arg] % R)
        else:
            raise ValueError("bad op %r" % (op,))
    assert len(st) == 1
    return st[0]


def blind(seed, obj, i):
    """The injected randomness: element i of stream `obj`, uniform in [0, r)."""
    import hashlib
    ctr = 0
    while True:
        h = hashlib.blake2b(b"h2agg-blind" + seed.to_bytes(8, "little") + obj.to_bytes(4, "little") + i.to_bytes(8, "little")
                            + ctr.to_bytes(4, "little"), digest_size=32).digest()
        v = int.from_bytes(h, "little") & ((1 << 254) - 1)
        if v < R:
            return v
        ctr += 1


# object ids of the blinding streams
def BL_LOOKUP_A(i): return 100 + 3 * i
def BL_LOOKUP_S(i): return 101 + 3 * i
def BL_LOOKUP_Z(i): return 102 + 3 * i
def BL_PERM_Z(i): return 1000 + i
BL_RANDOM_POLY = 5000


def blinds_buffer(shape, seed):
    """The prover's randomness in the layout of include/h2agg.h `h2a_create_proof` (canonical integers)."""
    n, u, bf = shape.n, shape.usable, shape.bf
    out = []
    for li in range(len(shape.lookups)):
        out += [blind(seed, BL_LOOKUP_A(li), i) for i in range(n - u)]
        out += [blind(seed, BL_LOOKUP_S(li), i) for i in range(n - u)]
        out += [blind(seed, BL_LOOKUP_Z(li), i) for i in range(bf)]
    cl = shape.chunk_len
    for ci in range((len(shape.perm_columns) + cl - 1) // cl):
        out += [blind(seed, BL_PERM_Z(ci), i) for i in range(bf)]
    out += [blind(seed, BL_RANDOM_POLY, i) for i in range(n)]
    return out


class Params:
    """KZG setup for tests: g[i] = [s^i]G, g_lagrange[i] = [L_i(s)]G.  `s` is kept so the
    final pairing equation can be checked as a discrete-log relation in G1."""

    def __init__(self, orc, k, s):
        import numpy as np
        self.k, self.n, self.s = k, 1 << k, s % R
        n = self.n
        omega = pm.omega_for(k)
        g = np.frombuffer(pm.affine_bytes(pm.G1), dtype=np.uint8)
        pw = [pow(self.s, i, R) for i in range(n)]
        # L_i(s) = omega^i (s^n - 1) / (n (s - omega^i))
        sn1 = (pow(self.s, n, R) - 1) % R
        lag = [pow(omega, i, R) * sn1 % R * inv(n * (self.s - pow(omega, i, R)) % R) % R for i in range(n)]
        self.g = np.concatenate([orc.g1_mul(g, np.frombuffer(pm.fr_mont_bytes(v), dtype=np.uint8)) for v in pw])
        self.g_lagrange = np.concatenate([orc.g1_mul(g, np.frombuffer(pm.fr_mont_bytes(v), dtype=np.uint8)) for v in lag])


def scalars_bytes(vals):
    import numpy as np
    return np.frombuffer(b"".join(pm.fr_mont_bytes(v) for v in vals), dtype=np.uint8)


def commit(orc, bases, vals):
    return pm.affine_from_bytes(orc.msm(bases[:64 * len(vals)], scalars_bytes(vals)))


# ------------------------------------------------------------------ keygen (fixed + permutation part of the keys)
def build_sigmas(shape, cycles):
    """sigma_j[row] = delta^j' * omega^row' of the cell (j', row') that follows (j, row) in its copy cycle.
    `cycles`: list of lists of (perm column position, row)."""
    n, omega = shape.n, shape.omega
    m = len(shape.perm_columns)
    nxt = {}
    for cyc in cycles:
        for a, b in zip(cyc, cyc[1:] + cyc[:1]):
            nxt[a] = b
    sig = []
    for j in range(m):
        col = []
        for row in range(n):
            jj, rr = nxt.get((j, row), (j, row))
            col.append(pow(DELTA, jj, R) * pow(omega, rr, R) % R)
        sig.append(col)
    return sig


class Keys:
    def __init__(self, orc, params, shape, fixed_values, sigmas, vk_hash):
        self.fixed_values, self.sigmas, self.vk_hash = fixed_values, sigmas, vk_hash % R
        self.fixed_commitments = [commit(orc, params.g_lagrange, col) for col in fixed_values]
        self.sigma_commitments = [commit(orc, params.g_lagrange, col) for col in sigmas]


# ------------------------------------------------------------------ shared expression list (prover rows / verifier point)
def permutation_and_lookup_expressions(shape, l0, llast, lblind, beta, gamma, theta, xpt, adv, fix, inst, sigma_evals,
                                       perm_sets, lookup_evals):
    """Expressions 1..4 of src/permutation.rs:211-321 then the five of src/lookup.rs:190-310, at one point.
    perm_sets: [(z, z_next, z_last_or_None)]; lookup_evals: [(z, z_next, a, a_prev, s)]."""
    out = []
    one_minus = (1 - (llast + lblind)) % R
    if perm_sets:
        out.append(l0 * (1 - perm_sets[0][0]) % R)
        zl = perm_sets[-1][0]
        out.append(llast * (zl * zl - zl) % R)
        for i in range(1, len(perm_sets)):
            out.append(l0 * (perm_sets[i][0] - perm_sets[i - 1][2]) % R)
        cl = shape.chunk_len
        for ci, (z, zn, _zl) in enumerate(perm_sets):
            cols = shape.perm_columns[ci * cl:(ci + 1) * cl]
            left, right = zn, z
            for i, (ctype, _cidx, qidx) in enumerate(cols):
                val = (adv, fix, inst)[ctype][qidx]
                left = left * ((beta * sigma_evals[ci * cl + i] + val + gamma) % R) % R
                right = right * ((beta * pow(DELTA, ci * cl + i, R) % R * xpt + val + gamma) % R) % R
            out.append((left - right) * one_minus % R)
    for (inputs, tables), (z, zn, a, aprev, s) in zip(shape.lookups, lookup_evals):
        out.append(l0 * (1 - z) % R)
        out.append(llast * (z * z - z) % R)
        ci = 0
        for prog in inputs:
            ci = (ci * theta + eval_rpn(prog, shape.constants, adv, fix, inst)) % R
        ct = 0
        for prog in tables:
            ct = (ct * theta + eval_rpn(prog, shape.constants, adv, fix, inst)) % R
        left = (a + beta) * (s + gamma) % R * zn % R
        right = (ci + beta) * (ct + gamma) % R * z % R
        out.append((left - right) * one_minus % R)
        out.append(l0 * (a - s) % R)
        out.append((a - s) * (a - aprev) % R * one_minus % R)
    return out


def rotate_point(shape, x, rot):
    return x * (pow(shape.omega, rot, R) if rot >= 0 else pow(inv(shape.omega), -rot, R)) % R


# ------------------------------------------------------------------ prover
def create_proof(orc, params, shape, keys, instance_cols, advice_cols, seed):
    """Returns (proof bytes, instance commitments).  advice_cols arrive already blinded (App. C)."""
    n, k, bf, u = shape.n, shape.k, shape.bf, shape.usable
    omega = shape.omega
    tr = pm.Blake2bTranscript()
    proof = bytearray()

    def write_point(p):
        tr.common_point(p); proof.extend(pm.compress_point(p))

    def write_scalar(s):
        tr.common_scalar(s); proof.extend(pm.le32(s))

    tr.common_scalar(keys.vk_hash)                                   # src/verifier.rs:341-358
    inst_comms = [commit(orc, params.g_lagrange, col) for col in instance_cols]
    for c in inst_comms:
        tr.common_point(c)                                           # :360-363
    for col in advice_cols:
        write_point(commit(orc, params.g_lagrange, col))             # :365-376
    theta = tr.squeeze_challenge()                                   # :378

    def row_queries(row):
        adv = [advice_cols[c][(row + r) % n] for c, r in shape.advice_queries]
        fix = [keys.fixed_values[c][(row + r) % n] for c, r in shape.fixed_queries]
        inst = [instance_cols[c][(row + r) % n] for c, r in shape.instance_queries]
        return adv, fix, inst

    lookups = []
    for li, (inputs, tables) in enumerate(shape.lookups):            # :380-387, src/lookup.rs:49-79
        A, S = [], []
        for row in range(n):
            adv, fix, inst = row_queries(row)
            ca = 0
            for prog in inputs:
                ca = (ca * theta + eval_rpn(prog, shape.constants, adv, fix, inst)) % R
            cs = 0
            for prog in tables:
                cs = (cs * theta + eval_rpn(prog, shape.constants, adv, fix, inst)) % R
            A.append(ca); S.append(cs)
        pa = sorted(A[:u])
        left = {}
        for v in S[:u]:
            left[v] = left.get(v, 0) + 1
        ps = [None] * u
        repeated = []
        for row, v in enumerate(pa):
            if row == 0 or v != pa[row - 1]:
                ps[row] = v
                assert left.get(v, 0) > 0, "lookup input not in table"
                left[v] -= 1
            else:
                repeated.append(row)
        for v in sorted(left):
            for _ in range(left[v]):
                ps[repeated.pop()] = v
        assert not repeated
        pa += [blind(seed, BL_LOOKUP_A(li), i) for i in range(n - u)]
        ps += [blind(seed, BL_LOOKUP_S(li), i) for i in range(n - u)]
        write_point(commit(orc, params.g_lagrange, pa))
        write_point(commit(orc, params.g_lagrange, ps))
        lookups.append(dict(A=A, S=S, pa=pa, ps=ps))
    beta = tr.squeeze_challenge()                                    # :390
    gamma = tr.squeeze_challenge()                                   # :393

    def column_values(ctype, cidx):
        return (advice_cols, keys.fixed_values, instance_cols)[ctype][cidx]

    perm_z = []
    last_z = 1
    cl = shape.chunk_len
    for ci in range((len(shape.perm_columns) + cl - 1) // cl):       # :402-409, src/permutation.rs:48-78
        cols = shape.perm_columns[ci * cl:(ci + 1) * cl]
        mod = [1] * n
        for i, (ctype, cidx, _q) in enumerate(cols):
            vals, sig = column_values(ctype, cidx), keys.sigmas[ci * cl + i]
            for row in range(n):
                mod[row] = mod[row] * ((beta * sig[row] + gamma + vals[row]) % R) % R
        mod = batch_inv(mod)
        for i, (ctype, cidx, _q) in enumerate(cols):
            vals = column_values(ctype, cidx)
            dw = pow(DELTA, ci * cl + i, R) * beta % R
            for row in range(n):
                mod[row] = mod[row] * ((dw + gamma + vals[row]) % R) % R
                dw = dw * omega % R
        z = [last_z]
        for row in range(1, n):
            z.append(z[-1] * mod[row - 1] % R)
        for i in range(bf):
            z[n - bf + i] = blind(seed, BL_PERM_Z(ci), i)
        last_z = z[u]
        write_point(commit(orc, params.g_lagrange, z))
        perm_z.append(z)
    for li, lk in enumerate(lookups):                                # :411-417, src/lookup.rs:81-106
        num = [(lk["A"][i] + beta) * (lk["S"][i] + gamma) % R for i in range(u)]
        den = batch_inv([(lk["pa"][i] + beta) * (lk["ps"][i] + gamma) % R for i in range(u)])
        z = [1]
        for i in range(u):
            z.append(z[-1] * num[i] % R * den[i] % R)
        assert z[u] == 1, "lookup grand product does not close"
        z += [blind(seed, BL_LOOKUP_Z(li), i) for i in range(bf)]
        assert len(z) == n
        write_point(commit(orc, params.g_lagrange, z))
        lk["z"] = z
    random_poly = [blind(seed, BL_RANDOM_POLY, i) for i in range(n)]
    write_point(commit(orc, params.g, random_poly))                  # :419-421, src/vanishing.rs:54-75
    y = tr.squeeze_challenge()                                       # :423

    # ---- quotient on the extended coset
    ext_n = 1 << shape.ext_k
    ext_omega = pm.omega_for(shape.ext_k)
    g = shape.coset_shift
    step = ext_n // n

    def to_coeff(vals):
        return intt(vals, omega)

    def to_ext(coeffs):
        return ntt([c * pow(g, i, R) % R for i, c in enumerate(coeffs)] + [0] * (ext_n - len(coeffs)), ext_omega)

    adv_c = [to_coeff(c) for c in advice_cols]
    fix_c = [to_coeff(c) for c in keys.fixed_values]
    inst_c = [to_coeff(c) for c in instance_cols]
    sig_c = [to_coeff(c) for c in keys.sigmas]
    pz_c = [to_coeff(z) for z in perm_z]
    for lk in lookups:
        lk["pa_c"], lk["ps_c"], lk["z_c"] = to_coeff(lk["pa"]), to_coeff(lk["ps"]), to_coeff(lk["z"])
    unit = lambda i: [1 if r == i else 0 for r in range(n)]
    l0_e = to_ext(to_coeff(unit(0)))
    llast_e = to_ext(to_coeff(unit(u)))
    lblind_e = to_ext(to_coeff([1 if r > u else 0 for r in range(n)]))
    adv_e, fix_e, inst_e = [to_ext(c) for c in adv_c], [to_ext(c) for c in fix_c], [to_ext(c) for c in inst_c]
    sig_e, pz_e = [to_ext(c) for c in sig_c], [to_ext(c) for c in pz_c]
    for lk in lookups:
        lk["pa_e"], lk["ps_e"], lk["z_e"] = to_ext(lk["pa_c"]), to_ext(lk["ps_c"]), to_ext(lk["z_c"])
    last_rot = -(bf + 1)
    h_ext = []
    for i in range(ext_n):
        at = lambda e, rot: e[(i + rot * step) % ext_n]
        adv = [at(adv_e[c], r) for c, r in shape.advice_queries]
        fix = [at(fix_e[c], r) for c, r in shape.fixed_queries]
        inst = [at(inst_e[c], r) for c, r in shape.instance_queries]
        xpt = g * pow(ext_omega, i, R) % R
        exprs = [eval_rpn(p, shape.constants, adv, fix, inst) for p in shape.gates]
        perm_sets = [(at(z, 0), at(z, 1), at(z, last_rot)) for z in pz_e]
        lk_evals = [(at(lk["z_e"], 0), at(lk["z_e"], 1), at(lk["pa_e"], 0), at(lk["pa_e"], -1), at(lk["ps_e"], 0)) for lk in lookups]
        exprs += permutation_and_lookup_expressions(shape, l0_e[i], llast_e[i], lblind_e[i], beta, gamma, theta, xpt, adv, fix, inst,
                                                    [s[i] for s in sig_e], perm_sets, lk_evals)
        acc = 0
        for e in exprs:
            acc = (acc * y + e) % R
        h_ext.append(acc * inv((pow(xpt, n, R) - 1) % R) % R)      # divide by the vanishing polynomial
    h_coeff = intt(h_ext, ext_omega)
    h_coeff = [c * pow(inv(g), i, R) % R for i, c in enumerate(h_coeff)]
    qd = shape.quotient_poly_degree
    assert all(c == 0 for c in h_coeff[qd * n:]), "quotient has too high a degree: constraints not satisfied"
    h_pieces = [h_coeff[i * n:(i + 1) * n] for i in range(qd)]
    for piece in h_pieces:
        write_point(commit(orc, params.g, piece))                    # :427-434, src/vanishing.rs:77-106
    x = tr.squeeze_challenge()                                       # :436

    # ---- evaluations, in the order _verify_proof reads them (:438-510)
    xr = lambda rot: rotate_point(shape, x, rot)
    inst_evals = [poly_eval(inst_c[c], xr(r)) for c, r in shape.instance_queries]
    adv_evals = [poly_eval(adv_c[c], xr(r)) for c, r in shape.advice_queries]
    fix_evals = [poly_eval(fix_c[c], xr(r)) for c, r in shape.fixed_queries]
    for e in inst_evals + adv_evals + fix_evals:
        write_scalar(e)
    random_eval = poly_eval(random_poly, x)
    write_scalar(random_eval)                                        # src/vanishing.rs:108-134
    sigma_evals = [poly_eval(c, x) for c in sig_c]
    for e in sigma_evals:
        write_scalar(e)                                              # src/permutation.rs:140-168
    perm_sets = []
    for i, c in enumerate(pz_c):                                     # src/permutation.rs:81-138
        z0, z1 = poly_eval(c, x), poly_eval(c, xr(1))
        zl = poly_eval(c, xr(last_rot)) if i + 1 < len(pz_c) else None
        write_scalar(z0); write_scalar(z1)
        if zl is not None:
            write_scalar(zl)
        perm_sets.append((z0, z1, zl))
    lk_evals = []
    for lk in lookups:                                               # src/lookup.rs:108-171
        ev = (poly_eval(lk["z_c"], x), poly_eval(lk["z_c"], xr(1)), poly_eval(lk["pa_c"], x), poly_eval(lk["pa_c"], xr(-1)),
              poly_eval(lk["ps_c"], x))
        for e in ev:
            write_scalar(e)
        lk_evals.append(ev)
    v = tr.squeeze_challenge()                                       # :718
    u_ch = tr.squeeze_challenge()                                    # :719  (the W_i are never absorbed)

    # ---- multi-open: the query list of :654-715 with polynomials in place of commitments
    xn = pow(x, n, R)
    h_poly = [0] * n
    xp = 1
    for piece in h_pieces:                                           # H = sum_i (x^n)^i h_i
        h_poly = [(a + xp * b) % R for a, b in zip(h_poly, piece)]
        xp = xp * xn % R
    h_eval = poly_eval(h_poly, x)
    queries = []
    queries += [(inst_c[c], r, e) for (c, r), e in zip(shape.instance_queries, inst_evals)]
    queries += [(adv_c[c], r, e) for (c, r), e in zip(shape.advice_queries, adv_evals)]
    for c, (z0, z1, _zl) in zip(pz_c, perm_sets):
        queries += [(c, 0, z0), (c, 1, z1)]
    for c, (_z0, _z1, zl) in list(zip(pz_c, perm_sets))[::-1][1:]:
        queries.append((c, last_rot, zl))
    for lk, ev in zip(lookups, lk_evals):
        queries += [(lk["z_c"], 0, ev[0]), (lk["pa_c"], 0, ev[2]), (lk["ps_c"], 0, ev[4]), (lk["pa_c"], -1, ev[3]), (lk["z_c"], 1, ev[1])]
    queries += [(fix_c[c], r, e) for (c, r), e in zip(shape.fixed_queries, fix_evals)]
    queries += [(c, 0, e) for c, e in zip(sig_c, sigma_evals)]
    queries += [(h_poly, 0, h_eval), (random_poly, 0, random_eval)]
    sets = {}
    for q in queries:
        sets.setdefault(q[1], []).append(q)
    for rot in sorted(sets):                                         # src/multiopen.rs:344-395 (prover mirror)
        pb, eb = [0] * n, 0
        for poly, _r, ev in sets[rot]:
            pb = [(a * v + b) % R for a, b in zip(pb, poly)]
            eb = (eb * v + ev) % R
        z = xr(rot)
        pb[0] = (pb[0] - eb) % R
        q = [0] * n                                                  # (pb)/(X - z), remainder must vanish
        carry = 0
        for i in range(n - 1, -1, -1):
            q[i] = carry
            carry = (pb[i] + carry * z) % R
        assert carry == 0, "opening is inconsistent"
        write_point(commit(orc, params.g, q))
    return bytes(proof), inst_comms


# ------------------------------------------------------------------ verifier (restates VerifierChip::_verify_proof natively)
def verify_proof(shape, keys_fixed_commitments, keys_sigma_commitments, vk_hash, inst_comms, proof):
    """Returns dict(e, f, w, zw, queries, ws, x, u, v, ...) — everything calc_witness needs and produces."""
    n, bf = shape.n, shape.bf
    tr = pm.Blake2bTranscript()
    pos = [0]

    def read_point():
        p = pm.decompress_point(proof[pos[0]:pos[0] + 32]); pos[0] += 32
        return p

    def read_scalar():
        s = pm.from_le(proof[pos[0]:pos[0] + 32]); pos[0] += 32
        assert s < R
        return s

    def point():
        p = read_point(); tr.common_point(p); return p

    def scalar():
        s = read_scalar(); tr.common_scalar(s); return s

    tr.common_scalar(vk_hash)
    for c in inst_comms:
        tr.common_point(c)
    adv_comms = [point() for _ in range(shape.num_advice)]
    theta = tr.squeeze_challenge()
    lk_perm = [(point(), point()) for _ in shape.lookups]
    beta = tr.squeeze_challenge()
    gamma = tr.squeeze_challenge()
    cl = shape.chunk_len
    n_chunks = (len(shape.perm_columns) + cl - 1) // cl
    pz_comms = [point() for _ in range(n_chunks)]
    lk_z = [point() for _ in shape.lookups]
    random_comm = point()
    y = tr.squeeze_challenge()
    h_comms = [point() for _ in range(shape.quotient_poly_degree)]
    x = tr.squeeze_challenge()
    inst_evals = [scalar() for _ in shape.instance_queries]
    adv_evals = [scalar() for _ in shape.advice_queries]
    fix_evals = [scalar() for _ in shape.fixed_queries]
    random_eval = scalar()
    sigma_evals = [scalar() for _ in shape.perm_columns]
    perm_sets = []
    for i in range(n_chunks):
        z0, z1 = read_scalar(), read_scalar()
        zl = read_scalar() if i + 1 < n_chunks else None
        tr.common_scalar(z0); tr.common_scalar(z1)
        if zl is not None:
            tr.common_scalar(zl)
        perm_sets.append((z0, z1, zl))
    lk_evals = [tuple(scalar() for _ in range(5)) for _ in shape.lookups]

    xn = pow(x, n, R)                                               # :513-516
    omega_inv = inv(shape.omega)
    l_evals, wp = [], 1                                             # :553-591
    for _ in range(2 + bf):
        l_evals.append(wp * (xn - 1) % R * inv(n * (x - wp) % R) % R)
        wp = wp * omega_inv % R
    l_evals.reverse()
    l_last, l_blind, l_0 = l_evals[0], sum(l_evals[1:1 + bf]) % R, l_evals[1 + bf]
    exprs = [eval_rpn(p, shape.constants, adv_evals, fix_evals, inst_evals) for p in shape.gates]
    exprs += permutation_and_lookup_expressions(shape, l_0, l_last, l_blind, beta, gamma, theta, x, adv_evals, fix_evals, inst_evals,
                                                sigma_evals, perm_sets, lk_evals)
    h_eval = exprs[0]                                               # src/vanishing.rs:145-175
    for e in exprs[1:]:
        h_eval = (h_eval * y + e) % R
    h_eval = h_eval * inv((xn - 1) % R) % R
    H = pm.msm([pow(xn, i, R) for i in range(len(h_comms))], h_comms)   # :177-188

    last_rot = -(bf + 1)
    queries = []                                                    # :654-715
    queries += [(inst_comms[c], r, e) for (c, r), e in zip(shape.instance_queries, inst_evals)]
    queries += [(adv_comms[c], r, e) for (c, r), e in zip(shape.advice_queries, adv_evals)]
    for c, (z0, z1, _zl) in zip(pz_comms, perm_sets):
        queries += [(c, 0, z0), (c, 1, z1)]
    for c, (_z0, _z1, zl) in list(zip(pz_comms, perm_sets))[::-1][1:]:
        queries.append((c, last_rot, zl))
    for (pa, ps), zc, ev in zip(lk_perm, lk_z, lk_evals):
        queries += [(zc, 0, ev[0]), (pa, 0, ev[2]), (ps, 0, ev[4]), (pa, -1, ev[3]), (zc, 1, ev[1])]
    queries += [(keys_fixed_commitments[c], r, e) for (c, r), e in zip(shape.fixed_queries, fix_evals)]
    queries += [(c, 0, e) for c, e in zip(keys_sigma_commitments, sigma_evals)]
    queries += [(H, 0, h_eval), (random_comm, 0, random_eval)]
    v = tr.squeeze_challenge()
    u = tr.squeeze_challenge()
    n_sets = len({q[1] for q in queries})
    ws = [read_point() for _ in range(n_sets)]
    assert pos[0] == len(proof), "trailing bytes in proof"
    e, f, w, zw = pm.gwc_accumulate(queries, ws, x, u, v, shape.omega)
    return dict(e=e, f=f, w=w, zw=zw, queries=queries, ws=ws, x=x, u=u, v=v, y=y, theta=theta, beta=beta, gamma=gamma,
                h_eval=h_eval, H=H)


def pairing_relation_holds(res, s):
    """e(W, [s]_2) == e(ZW + F + E, [1]_2), checked in G1 with the known setup secret."""
    lhs = pm.g1_mul(res["w"], s)
    rhs = pm.g1_add(pm.g1_add(res["zw"], res["f"]), res["e"])
    return lhs == rhs
