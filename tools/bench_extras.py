"""The blocks bench.py prints next to its headline: strong-scaling MSMs, NTT / coset extension, the prover pipeline,
a batch of proofs, and the cold (upload + table build + first MSM) figure.  Every block is gated on the CPU oracle
that bench.py passes in (`env["orc"]`, `env["pk"]`): this module never imports `oracle/` itself.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

MODMUL_PER_ADD = 6


def _oracle_sum(orc, pts):
    acc = np.zeros(64, np.uint8)
    for i in range(pts.size // 64):
        acc = orc.g1_add(acc, pts[64 * i:64 * i + 64])
    return acc


def _msm_case(env, seed_b, seed_s, total, steps, tag):
    """One `total`-point MSM split into point ranges over the ranks; device-resident, tables built once."""
    ctx, h2a, orc, torch, world, rank = (env[k] for k in ("ctx", "h2a", "orc", "torch", "world", "rank"))
    per = total // world
    d_b = torch.empty(64 * per, dtype=torch.uint8, device="cuda")
    d_s = torch.empty(32 * per, dtype=torch.uint8, device="cuda")
    ctx.gen_bases_dev(seed_b, per, d_b.data_ptr(), first=rank * per)
    ctx.gen_scalars_dev(seed_s, per, d_s.data_ptr(), first=rank * per)
    hb = ctx.bases_from_device(d_b.data_ptr(), per)
    if not env["args"].no_precompute:
        hb.precompute(-1)
    step = lambda: env["combine"](ctx.msm_dev(hb, d_s.data_ptr(), per))
    for _ in range(3):
        step()
    ms, _, launches, phases, result = env["timed"](step, steps)
    mine = ctx.msm_dev(hb, d_s.data_ptr(), per)
    t0 = time.perf_counter()
    want = orc.msm(d_b.cpu().numpy(), d_s.cpu().numpy(), threads=env["cpu_threads"])
    cpu_s = time.perf_counter() - t0
    ok = bytes(mine) == bytes(want)
    if world > 1:
        ok = ok and bytes(_oracle_sum(orc, h2a.allgather_points(want, device="cuda"))) == bytes(result)
    if not env["all_ranks_ok"](ok):
        raise SystemExit("bench.py: PARITY FAILURE — strong-scaling MSM (%s) differs from the oracle" % tag)
    acc = sum(ph[3][1] for ph in phases) / len(phases)
    windows = 13 if not env["args"].no_precompute else 16
    out = {"total_points": total, "points_per_rank": per, "ms_per_step": ms / steps, "value": total / (ms / steps * 1e-3) / 1e6,
           "unit": "Mpts/s", "steps": steps, "parity_checked": "oracle",
           "int_pipe_frac": (per * windows * MODMUL_PER_ADD / (acc * 1e-3) / 1e9 / env["modmul_peak"]) if acc else None,
           "accumulate_ms": acc}
    if world == 1:
        out["cpu_baseline"] = {"value": total / cpu_s / 1e6, "unit": "Mpts/s", "cores": env["cpu_threads"], "kind": "port",
                               "sample": "one best_multiexp over all %d points, %.2f s" % (total, cpu_s)}
    hb.free()
    del d_b, d_s
    torch.cuda.empty_cache()
    return out


def strong_block(env, headline):
    """A fixed 2^22-point and a fixed 2^24-point MSM over N ranks (point ranges, 64-byte allgather)."""
    bases, d_scal, n, value, ms_per_step, result = headline
    world, args = env["world"], env["args"]
    steps = min(args.steps, 10)
    out = {"scaling": "strong", "note": "total size fixed, each rank takes total/N consecutive points; value = total points / max-over-ranks device time"}
    if world == 1 and args.log_n == 22:
        out["2^22"] = {"total_points": n, "points_per_rank": n, "ms_per_step": ms_per_step, "value": value, "unit": "Mpts/s",
                       "parity_checked": "oracle", "note": "the headline measurement (N = 1)"}
    else:
        out["2^22"] = _msm_case(env, 1, 2, 1 << 22, steps, "2^22")
    out["2^24"] = _msm_case(env, 5, 6, 1 << 24, max(3, steps // 2), "2^24")
    return out


def cold_block(env, d_bases, host_scalars, n, want):
    """Cold start: bases from (pageable) host memory -> HBM, window tables built, first MSM from host scalars."""
    ctx, torch = env["ctx"], env["torch"]
    host_bases = d_bases.cpu().numpy()
    ctx.sync()
    t0 = time.perf_counter()
    hb = ctx.upload_bases(host_bases)
    t1 = time.perf_counter()
    if not env["args"].no_precompute:
        hb.precompute(-1)
    ctx.sync()
    t2 = time.perf_counter()
    got = ctx.msm(hb, host_scalars)
    t3 = time.perf_counter()
    hb.free()
    if bytes(got) != bytes(want):
        raise SystemExit("bench.py: cold-start MSM differs from the resident-bases result")
    return {"total_s": t3 - t0, "upload_bases_s": t1 - t0, "table_build_s": t2 - t1, "first_msm_s": t3 - t2,
            "value": n / (t3 - t0) / 1e6, "unit": "Mpts/s", "h2d_bytes": 96 * n,
            "note": "what one MSM costs when nothing is resident: 64 B/point bases upload from pageable host memory + h2a_bases_precompute + "
                    "h2a_msm_g1 with host scalars; the tables pay for themselves after table_build_s / (no-table ms - table ms) MSMs over the same Params"}


def ntt_block(env):
    """Fr NTT at k = 22 (device-resident and through host buffers) and the k=20 -> 2^22 coset extension of the quotient path."""
    ctx, h2a, orc, torch = env["ctx"], env["h2a"], env["orc"], env["torch"]
    args = env["args"]
    steps = min(args.steps, 10)
    out = {}
    for tag, k, ext_k in (("forward k=22", 22, 0), ("coeff_to_extended 2^20 -> 2^22", 20, 22), ("forward k=24", 24, 0)):
        log_n = ext_k or k
        n_in, n = 1 << k, 1 << log_n
        d_in = torch.empty(32 * n_in, dtype=torch.uint8, device="cuda")
        d_out = torch.empty(32 * n, dtype=torch.uint8, device="cuda")
        ctx.gen_scalars_dev(7, n_in, d_in.data_ptr())
        host_in = d_in.cpu().numpy()
        omega = h2a.fr_root_of_unity(log_n)
        shift = ctx.field_op(1, "to_mont", np.frombuffer((7).to_bytes(32, "little"), dtype=np.uint8))
        if ext_k:
            run = lambda: ctx.coeff_to_extended_dev(d_in.data_ptr(), k, ext_k, shift, d_out.data_ptr())
        else:
            def run():
                d_out.copy_(d_in)       # on torch's stream; ordered before the transform by the barrier-free sync below
                torch.cuda.current_stream().synchronize()
                ctx.ntt_dev(d_out.data_ptr(), k, omega)
        for _ in range(3):
            run()
        phases = []
        for _ in range(steps):
            run()
            phases.append(ctx.last_phases(1))
        kernel_ms = sum(sum(ms for _, ms in ph) for ph in phases) / len(phases)
        got = d_out.cpu().numpy()
        t0 = time.perf_counter()
        want = orc.coeff_to_extended(host_in, k, ext_k, shift) if ext_k else orc.fft(host_in, k, omega)
        cpu_s = time.perf_counter() - t0
        if bytes(got) != bytes(want):
            raise SystemExit("bench.py: PARITY FAILURE — NTT (%s) differs from the oracle" % tag)
        products = (n // 2) * log_n
        entry = {"n": n, "kernel_ms": kernel_ms, "passes_ms": [ms for _, ms in phases[-1]], "value": n / (kernel_ms * 1e-3) / 1e6,
                 "unit": "Melements/s", "parity_checked": "oracle",
                 "roofline": {"bound": "hbm", "achieved": 64 * n / (kernel_ms * 1e-3) / 1e9, "peak": env["hbm_peak"], "unit": "GB/s",
                              "frac": 64 * n / (kernel_ms * 1e-3) / 1e9 / env["hbm_peak"], "peak_source": env["peak_src"],
                              "note": "algorithmic 64 B per element (one read + one write) over the sum of the pass kernels"},
                 "int_pipe": {"achieved": products / (kernel_ms * 1e-3) / 1e9, "peak": env["modmul_peak"], "unit": "1e9 Montgomery products/s",
                              "frac": products / (kernel_ms * 1e-3) / 1e9 / env["modmul_peak"],
                              "algorithmic": "(n/2) log2 n = %d butterfly products" % products},
                 "cpu_baseline": {"value": n / cpu_s / 1e6, "unit": "Melements/s", "cores": env["cpu_threads"], "kind": "port",
                                  "sample": "one best_fft (oracle/ restatement) of the same input, %.2f s" % cpu_s}}
        if not ext_k and k == 22:      # through the host-pointer entry point: H2D + transform + D2H
            pinned = torch.from_numpy(host_in.copy()).pin_memory().numpy()
            src = pinned.copy()
            ctx.ntt(pinned, k, omega, inplace=True)
            dt = 0.0
            for _ in range(steps):
                pinned[:] = src
                t0 = time.perf_counter()
                ctx.ntt(pinned, k, omega, inplace=True)
                dt += (time.perf_counter() - t0) / steps
            if bytes(pinned) != bytes(want):
                raise SystemExit("bench.py: PARITY FAILURE — host-buffer NTT differs from the oracle")
            entry["e2e"] = {"value": n / dt / 1e6, "unit": "Melements/s", "ms_per_step": dt * 1e3, "h2d_bytes_per_step": 32 * n,
                            "d2h_bytes_per_step": 32 * n}
        out[tag] = entry
        del d_in, d_out
        torch.cuda.empty_cache()
    return out


def _pk_shape(pk, s):
    return pk.Shape(k=s.k, blinding_factors=s.bf, degree=s.degree, num_instance=s.num_instance, num_advice=s.num_advice,
                    num_fixed=s.num_fixed, advice_queries=s.advice_queries, fixed_queries=s.fixed_queries,
                    instance_queries=s.instance_queries, gates=s.gates, constants=s.constants, lookups=s.lookups,
                    perm_columns=s.perm_columns, coset_shift=7)


def _oracle_accepts(env, chk):
    """The oracle's verifier (restating VerifierChip::_verify_proof) replays the proof and the pairing relation must hold."""
    pk, pm = env["pk"], env["pm"]
    to_pts = lambda b: [pm.affine_from_bytes(bytes(b[64 * i:64 * i + 64])) for i in range(len(b) // 64)]
    res = pk.verify_proof(_pk_shape(pk, chk["shape"]), to_pts(chk["fixed_commitments"]), to_pts(chk["sigma_commitments"]), chk["vk_hash"],
                          to_pts(chk["inst"]), bytes(chk["proof"]))
    return pk.pairing_relation_holds(res, chk["secret"]), res


def sweep_block(env):
    """BASELINE configs 2 and 3 below the sizes the other blocks cover: G1 MSM 2^16..2^20 and Fr NTT k = 16..20, inputs resident
    in HBM, each checked against the oracle and timed beside it (the 2^22 / 2^24 points of both sweeps are `strong` and `ntt`)."""
    ctx, h2a, orc, torch = env["ctx"], env["h2a"], env["orc"], env["torch"]
    out = {"msm": {}, "ntt": {}}
    for log_n in (16, 18, 20):
        n = 1 << log_n
        d_b = torch.empty(64 * n, dtype=torch.uint8, device="cuda")
        d_s = torch.empty(32 * n, dtype=torch.uint8, device="cuda")
        ctx.gen_bases_dev(21, n, d_b.data_ptr())
        ctx.gen_scalars_dev(22, n, d_s.data_ptr())
        hb = ctx.bases_from_device(d_b.data_ptr(), n)
        if not env["args"].no_precompute:
            hb.precompute(-1)
        step = lambda: ctx.msm_dev(hb, d_s.data_ptr(), n)
        for _ in range(3):
            step()
        ms, _, _, _, got = env["timed"](step, 10)
        t0 = time.perf_counter()
        want = orc.msm(d_b.cpu().numpy(), d_s.cpu().numpy(), threads=env["cpu_threads"])
        cpu_s = time.perf_counter() - t0
        if bytes(got) != bytes(want):
            raise SystemExit("bench.py: PARITY FAILURE — MSM sweep 2^%d differs from the oracle" % log_n)
        out["msm"]["2^%d" % log_n] = {"ms": ms / 10, "value": n / (ms / 10 * 1e-3) / 1e6, "unit": "Mpts/s", "parity_checked": "oracle",
                                      "cpu_baseline": {"value": n / cpu_s / 1e6, "unit": "Mpts/s", "cores": env["cpu_threads"], "kind": "port",
                                                       "sample": "one best_multiexp over the same points, %.3f s" % cpu_s}}
        hb.free()
        del d_b, d_s
    for k in (16, 18, 20):
        n = 1 << k
        d_in = torch.empty(32 * n, dtype=torch.uint8, device="cuda")
        d_out = torch.empty(32 * n, dtype=torch.uint8, device="cuda")
        ctx.gen_scalars_dev(23, n, d_in.data_ptr())
        omega = h2a.fr_root_of_unity(k)
        phases = []
        for it in range(8):
            d_out.copy_(d_in)
            torch.cuda.current_stream().synchronize()
            ctx.ntt_dev(d_out.data_ptr(), k, omega)
            if it >= 3:
                phases.append(ctx.last_phases(1))
        kernel_ms = sum(sum(ms for _, ms in ph) for ph in phases) / len(phases)
        host_in = d_in.cpu().numpy()
        t0 = time.perf_counter()
        want = orc.fft(host_in, k, omega)
        cpu_s = time.perf_counter() - t0
        if bytes(d_out.cpu().numpy()) != bytes(want):
            raise SystemExit("bench.py: PARITY FAILURE — NTT sweep k=%d differs from the oracle" % k)
        out["ntt"]["k=%d" % k] = {"kernel_ms": kernel_ms, "value": n / (kernel_ms * 1e-3) / 1e6, "unit": "Melements/s", "parity_checked": "oracle",
                                  "cpu_baseline": {"value": n / cpu_s / 1e6, "unit": "Melements/s", "cores": env["cpu_threads"], "kind": "port",
                                                   "sample": "one best_fft of the same input, %.3f s" % cpu_s}}
        del d_in, d_out
    torch.cuda.empty_cache()
    return out


def mulvar_block(env):
    """Row f4: witness cells of the non-native mul_var of 64 aggregated proofs (37 each) at once, 64 entries checked cell for cell
    against the compiled oracle (which is also the CPU figure beside it) and one against the big-integer statement (env["mv"])."""
    ctx, pm, mv, torch = env["ctx"], env["pm"], env["mv"], env["torch"]
    m = 37 * 64
    ln = ctx.mulvar_witness_len()
    aux_pt = pm.g1_mul(pm.G1, 0xabcdef123457)
    aux = np.frombuffer(pm.affine_bytes(aux_pt), dtype=np.uint8)
    d_p = torch.empty(64 * m, dtype=torch.uint8, device="cuda")
    d_s = torch.empty(32 * m, dtype=torch.uint8, device="cuda")
    d_r = torch.empty(64 * m, dtype=torch.uint8, device="cuda")
    d_w = torch.empty(32 * ln * m, dtype=torch.uint8, device="cuda")
    ctx.gen_bases_dev(31, m, d_p.data_ptr())
    ctx.gen_scalars_dev(32, m, d_s.data_ptr())
    best = None
    for _ in range(3):
        ctx.sync()
        t0 = time.perf_counter()
        ctx.mulvar_witness_dev(d_p.data_ptr(), d_s.data_ptr(), m, aux, d_r.data_ptr(), d_w.data_ptr())
        ctx.sync()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    # parity: 64 entries spread over the batch, cell for cell against the compiled oracle (itself pinned to the big-integer
    # statement oracle/mulvar.py by tests/test_mulvar_oracle.py); the same call, over all host threads, is the CPU figure
    orc = env["orc"]
    pts, scal, res = d_p.cpu().numpy(), d_s.cpu().numpy(), d_r.cpu().numpy()
    idx = list(range(0, m, m // 64))[:64]
    sp = np.concatenate([pts[64 * i:64 * i + 64] for i in idx])
    ss = np.concatenate([scal[32 * i:32 * i + 32] for i in idx])
    t0 = time.perf_counter()
    want_res, want_cells, want_st = orc.mulvar_witness(sp, ss, aux, threads=env["cpu_threads"])
    cpu_s = time.perf_counter() - t0
    for j, i in enumerate(idx):
        got = d_w[32 * ln * i:32 * ln * (i + 1)].cpu().numpy()
        if want_st[j] != 0 or bytes(got) != bytes(want_cells[32 * ln * j:32 * ln * (j + 1)]) or bytes(res[64 * i:64 * i + 64]) != bytes(want_res[64 * j:64 * j + 64]):
            raise SystemExit("bench.py: PARITY FAILURE — mul_var witness cells differ from the oracle (entry %d)" % i)
    # and one entry against the big-integer statement itself
    q, cells, st = mv.mulvar_witness(pm.affine_from_bytes(bytes(pts[:64])), pm.fr_from_mont_bytes(bytes(scal[:32])), aux_pt)
    if st != 0 or bytes(d_w[:32 * ln].cpu().numpy()) != b"".join(pm.fr_mont_bytes(c) for c in cells):
        raise SystemExit("bench.py: PARITY FAILURE — mul_var witness cells differ from oracle/mulvar.py (entry 0)")
    del d_p, d_s, d_r, d_w
    torch.cuda.empty_cache()
    return {"metric": "mul_var witness generation, 64 proofs x 37 mul_var", "value": best, "unit": "s", "higher_is_better": False, "mul_var": m,
            "cells_per_mul_var": ln, "witness_bytes": 32 * ln * m, "write_gb_per_s": 32 * ln * m / best / 1e9,
            "parity_checked": "oracle: 64 entries cell for cell against the compiled restatement (oracle/oracle.cpp), one against the big-integer one "
                              "(oracle/mulvar.py); PARITY UNPINNED at the dependency boundary (halo2wrong's cell layout is not in the reference)",
            "cpu_baseline": {"value": cpu_s / len(idx) * m, "unit": "s", "cores": env["cpu_threads"], "kind": "port",
                             "sample": "%d of the %d mul_var on the compiled oracle over all host threads (%.2f s), scaled; sequential-style witness code: "
                                       "one Fermat inversion per affine formula" % (len(idx), m, cpu_s)}}


def params_block(env):
    """Row f3: the k=20 parameters written to and read back from a parameter file (both encodings), points compared."""
    import tempfile
    ctx = env["ctx"]
    import prove_bench
    k = 20
    g, gl = ctx.kzg_setup(k, prove_bench.fr(ctx, [0x1234567]))
    out = {"k": k}
    ref = g.download()
    with tempfile.TemporaryDirectory() as td:
        for compressed in (False, True):
            path = os.path.join(td, "halo2-%d.params" % k)
            t0 = time.perf_counter(); ctx.params_write(path, k, g, gl, compressed=compressed); tw = time.perf_counter() - t0
            size = os.path.getsize(path)
            t0 = time.perf_counter(); k2, g2, gl2, _ = ctx.params_read(path); tr = time.perf_counter() - t0
            if k2 != k or bytes(g2.download()) != bytes(ref):
                raise SystemExit("bench.py: parameter file round trip changed the points")
            g2.free(); gl2.free()
            os.remove(path)
            out["compressed" if compressed else "in_memory_form"] = {"file_bytes": size, "write_s": tw, "read_s": tr,
                                                                     "read_gb_per_s": size / tr / 1e9}
    g.free(); gl.free()
    out["note"] = "read = file -> pinned staging -> HBM with the Blake2b digest and the on-curve check of every point (square roots when compressed)"
    return out


def prove_roofline(op_counts, k, ext_k, table_bits, seconds, modmul_peak, hbm_peak, peak_src):
    """SURVEY §8d for the prove metric: the algorithmic bytes and Montgomery products of the MSMs and transforms of one proof
    (MSM: 96 B and ceil(254 / c) batched-affine additions of 6 products per point; NTT: 64 B per element and (m / 2) log2 m
    butterfly products), summed over op_counts, over the wall time of the proof.  Pure arithmetic (tests/test_bench_helpers.py)."""
    n, m = 1 << k, 1 << ext_k
    windows = (254 + table_bits - 1) // table_bits
    msm_products = op_counts["msm_n"] * n * windows * MODMUL_PER_ADD
    ntt_products = op_counts["ifft_n"] * (n // 2) * k + (op_counts["coset_fft_4n"] + op_counts["ifft_4n"]) * (m // 2) * ext_k
    nbytes = op_counts["msm_n"] * n * 96 + op_counts["ifft_n"] * n * 64 + (op_counts["coset_fft_4n"] + op_counts["ifft_4n"]) * m * 64
    gbs = nbytes / seconds / 1e9
    gmul = (msm_products + ntt_products) / seconds / 1e9
    return ({"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak if hbm_peak else None, "traffic": None,
             "peak_source": peak_src, "algorithmic_bytes": nbytes,
             "note": "96 B per MSM point + 64 B per transformed element, each once, over the wall time of one proof; the pipeline is bound by the "
                     "integer pipe, see int_pipe"},
            {"achieved": gmul, "peak": modmul_peak, "unit": "1e9 Montgomery products/s", "frac": gmul / modmul_peak if modmul_peak else None,
             "algorithmic": "%d MSM products (%d windows of %d bits) + %d butterfly products" % (msm_products, windows, table_bits, ntt_products),
             "note": "an UPPER estimate of the pipe's use: advice columns of the aggregation profile hold range-checked small values whose "
                     "upper window digits are zero and are skipped, and quotient / grand-product / evaluation products are not counted; the "
                     "figure for full-width scalars is the headline's int_pipe"})


def prove_block(env):
    """`agg-circuit prove s at k=20`: the prover pipeline at every N, the proof checked by the oracle's verifier."""
    import prove_bench
    ctx, orc, torch, world, rank, args = (env[k] for k in ("ctx", "orc", "torch", "world", "rank", "args"))

    class _A:
        k, steps, lookups, precompute = args.prove_k, 2, 9, prove_bench.PROVER_TABLE_BITS
    _A.world, _A.rank = world, rank
    pr = prove_bench.run(ctx, _A)
    chk = pr.pop("_check")
    ok, _ = _oracle_accepts(env, chk) if rank == 0 else (True, None)
    ok = ok and pr["proof_verifies"]
    value = pr["value"]
    if world > 1:
        import hashlib
        dist = env["dist"]
        t = torch.tensor([value], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        value = float(t[0])
        digest = torch.tensor(list(hashlib.sha256(bytes(chk["proof"])).digest()), dtype=torch.uint8, device="cuda")
        alld = [torch.empty_like(digest) for _ in range(world)]
        dist.all_gather(alld, digest)
        ok = ok and all(bool((d == alld[0]).all()) for d in alld)
    if not env["all_ranks_ok"](ok):
        raise SystemExit("bench.py: PARITY FAILURE — the k=%d proof is rejected by the oracle verifier (or differs across ranks)" % args.prove_k)
    out = {"metric": pr["metric"], "value": value, "unit": "s", "higher_is_better": False, "n_gpus": world,
           "parity_checked": "oracle verifier (oracle/plonk.py verify_proof + pairing relation)" + ("; byte-identical proof on every rank" if world > 1 else ""),
           "proof_bytes": pr["proof_bytes"], "phases_ms": pr["phases_ms"], "workload": pr["config"]["workload"],
           "msm_tables": pr["config"]["msm_tables"], "op_counts": pr["op_counts"],
           "distribution": "single GPU" if world == 1 else
                           "one proof over %d GPUs through the library's NCCL communicators: commitments, lookups and transforms column-parallel "
                           "(coefficient forms broadcast, extended forms sent as row windows), quotient row-parallel; permutation grand products, "
                           "evaluations and openings replicated" % world,
           "nvlink_traffic_this_rank": pr.get("nvlink_traffic_this_rank")}
    try:   # roofline fractions of the whole proof (algorithmic figures of SURVEY §8d); never worth losing the block for
        ext_k = args.prove_k + 2    # degree 5: the quotient lives on 4n points
        out["roofline"], out["int_pipe"] = prove_roofline(pr["op_counts"], args.prove_k, ext_k, int(pr["config"]["msm_tables"]), value,
                                                           env["modmul_peak"] * world, env["hbm_peak"] * world, env["peak_src"] + (" x %d GPUs" % world if world > 1 else ""))
    except Exception as e:   # noqa: BLE001
        out["roofline"] = {"error": "%s: %s" % (type(e).__name__, e)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # CPU figure: the MSM / FFT sequence of create_proof replayed with the oracle's best_multiexp / best_fft
        k = args.prove_k
        n = 1 << k
        th = env["cpu_threads"]
        cnt = pr["op_counts"]
        d_b = ctx.dev_alloc(64 * n)
        ctx.gen_bases_dev(1, n, d_b)
        bases, scal = ctx.d2h(d_b, 64 * n), orc.gen_scalars(2, n)
        ctx.dev_free(d_b)
        t0 = time.perf_counter(); orc.msm(bases, scal, threads=th); t_msm = time.perf_counter() - t0
        w = orc.fr_root_of_unity(k)
        t0 = time.perf_counter(); orc.ifft(scal, k, w, threads=th); t_ifft = time.perf_counter() - t0
        shift = orc.to_mont(1, np.frombuffer((7).to_bytes(32, "little"), dtype=np.uint8))
        t0 = time.perf_counter(); ext = orc.coeff_to_extended(scal, k, k + 2, shift, threads=th); t_ext = time.perf_counter() - t0
        t0 = time.perf_counter(); orc.extended_to_coeff(ext, k + 2, shift, threads=th); t_back = time.perf_counter() - t0
        total = cnt["msm_n"] * t_msm + cnt["ifft_n"] * t_ifft + cnt["coset_fft_4n"] * t_ext + cnt["ifft_4n"] * t_back
        out["cpu_baseline"] = {"value": total, "unit": "s", "cores": th, "kind": "port",
                               "sample": "one best_multiexp(2^%d) %.2f s, one ifft %.2f s, one coeff_to_extended(4n) %.2f s, one extended_to_coeff %.2f s, "
                                         "each timed once on the oracle port and multiplied by op_counts; omits quotient evaluation, grand products, "
                                         "lookup permutation, evaluations, Kate division and witness synthesis, so it UNDERSTATES the CPU prover" % (k, t_msm, t_ifft, t_ext, t_back)}
    return out


def batch_block(env):
    """BASELINE config 5: 64 proofs, one share per rank: prove, verify-accumulate the share in one launch, allgather (e,f,w,zw)."""
    import prove_bench
    ctx, h2a, torch, world, rank = (env[k] for k in ("ctx", "h2a", "torch", "world", "rank"))
    n_proofs, k = 64, 12
    lo, hi = h2a.shard_range(n_proofs, rank, world)
    secret = 0x0f1e2d3c4b5a69788796a5b4c3d2e1f00112233445566778899aabbccddeeff % prove_bench.R
    g, gl = ctx.kzg_setup(k, prove_bench.fr(ctx, [secret]))
    g.precompute(-1); gl.precompute(-1)
    shape, inst_b, adv_b, fixed_b, sigmas_b = prove_bench.build(ctx, k, n_lookups=1)
    circ = h2a.Circuit(ctx, shape, np.zeros(0, np.uint8))
    circ.set_keys(g, gl, fixed_b, sigmas_b, prove_bench.fr(ctx, [0xC0FFEE]), prove_bench.fr(ctx, [7]))
    circ.prove(inst_b, adv_b, prove_bench.random_blinds(ctx, circ.blinds_len(), 999))   # warm-up
    blinds = [prove_bench.random_blinds(ctx, circ.blinds_len(), p) for p in range(lo, hi)]
    env["barrier"]()
    t0 = time.perf_counter()
    proofs, insts = [], []
    for b in blinds:
        proof, inst = circ.prove(inst_b, adv_b, b)
        proofs.append(proof); insts.append(inst)
    ctx.sync()
    env["barrier"]()
    t_prove = time.perf_counter() - t0
    circ.verify_batch(np.concatenate(insts), proofs)       # warm-up
    env["barrier"]()
    t0 = time.perf_counter()
    mine = circ.verify_batch(np.concatenate(insts), proofs).reshape(-1)
    allr = h2a.allgather_points(mine, device="cuda") if world > 1 else mine
    env["barrier"]()
    t_verify = time.perf_counter() - t0
    # every (e, f, w, zw) of this rank's share must satisfy s*W == ZW + F + E; proof `lo` also goes through the oracle verifier
    ok = allr.size == 256 * n_proofs
    sfr = prove_bench.fr(ctx, [secret])
    for i in range(hi - lo):
        e, f, w, zw = (mine[256 * i + 64 * j:256 * i + 64 * j + 64] for j in range(4))
        ok = ok and bytes(ctx.msm_adhoc(w, sfr)) == bytes(h2a.g1_sum(np.concatenate([zw, f, e])))
    if hi > lo:
        fc, sc = circ.get_vk(shape.num_fixed, len(shape.perm_columns))
        good, res = _oracle_accepts(env, dict(shape=shape, fixed_commitments=fc, sigma_commitments=sc, vk_hash=0xC0FFEE, inst=insts[0],
                                              proof=proofs[0], secret=secret))
        pm = env["pm"]
        want = b"".join(pm.affine_bytes(res[nm]) for nm in ("e", "f", "w", "zw"))
        ok = ok and good and bytes(mine[:256]) == want
    if not env["all_ranks_ok"](ok):
        raise SystemExit("bench.py: PARITY FAILURE — batch verify-accumulate disagrees with the oracle verifier")
    t = torch.tensor([t_prove, t_verify], dtype=torch.float64, device="cuda")
    if world > 1:
        env["dist"].all_reduce(t, op=env["dist"].ReduceOp.MAX)
    circ.free(); g.free(); gl.free()
    return {"proofs": n_proofs, "k": k, "proofs_per_rank": hi - lo, "prove_s": float(t[0]), "verify_accumulate_s": float(t[1]),
            "prove_proofs_per_s": n_proofs / float(t[0]), "verify_proofs_per_s": n_proofs / float(t[1]),
            "parity_checked": "oracle verifier on one proof per rank + pairing relation on every (e,f,w,zw)",
            "note": "independent proofs, one share per GPU; 256-byte results allgathered as raw bytes"}
