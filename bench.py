#!/usr/bin/env python
"""bench.py — G1 MSM throughput at 2^22 points per GPU (BASELINE.json metric), one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--log-n 22]

A step is one MSM over this rank's 2^22-point range of a (2^22 * N)-point MSM; with N > 1 the 64-byte
per-rank partial results are allgathered (NCCL) and summed on every rank.  `value` times the step with
scalars and bases resident in HBM; `e2e` times the C-ABI call with HOST scalars (pinned), i.e. with the
host->device copy of the scalars and the device->host read of the result inside the timed region (bases
stay resident, as `Params` do across the commitments of one proof).

Parity gate: before a line is printed the timed result is compared with the CPU oracle (`oracle/`, the restatement of
halo2 `best_multiexp` / `best_fft`) on the same inputs, at the benchmark's own size; a mismatch aborts the run.

Next to the headline the line carries, each with its own roofline fractions and same-run CPU figure:
  `strong`   one fixed 2^22-point and one fixed 2^24-point MSM split over the N ranks (point ranges),
  `ntt`      Fr NTT at k=22 / k=24 and the k=20 -> 2^22 coset extension (N = 1 only: a transform does not shard),
  `sweep`    the small end of BASELINE configs 2 and 3: MSM 2^16..2^20 and NTT k=16..20, each beside the oracle (N = 1),
  `mulvar`   row f4: witness cells of the non-native mul_var of 64 aggregated proofs (N = 1),
  `params`   row f3: the k=20 parameter file written and read back, both encodings (N = 1),
  `prove`    the k=20 prover pipeline at every N (one proof spread over the ranks by the library's own NCCL plumbing:
             commitments, lookups and transforms column-parallel, quotient row-parallel),
  `batch`    64 proofs proved + verify-accumulated, one share per rank (BASELINE config 5),
  `e2e_cold` bases upload + window-table build + first MSM from host scalars.
`--impl reference` times the CPU restatement of the reference's path (oracle/, halo2 `best_multiexp`)
on the host cores instead.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "G1 MSM Mpts/s at 2^22"
UNIT = "Mpts/s"
ALGO_BYTES_PER_POINT = 96          # 64 B affine base + 32 B scalar, each read once
MODMUL_PER_ADD = 6                 # batched affine addition: 3 (Montgomery's trick) + 2M + 1S
NTT_BYTES_PER_ELEMENT = 64         # one read + one write of every element for the whole transform


def clocks_monitor_start(dev_index):
    q = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    try:
        return subprocess.Popen(["nvidia-smi", "-i", str(dev_index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "20"],
                                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except OSError:
        return None


def clocks_monitor_stop(proc, t_begin, t_end):
    """Samples whose timestamp falls inside the timed region [t_begin, t_end] (epoch seconds); when the region is
    too short to hold one, the samples taken under load since the monitor started (warm-up + timed region)."""
    if proc is None:
        return None
    import datetime
    proc.terminate()
    try:
        out, _ = proc.communicate(timeout=5)
    except subprocess.TimeoutExpired:
        proc.kill()
        out, _ = proc.communicate()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    rows = []
    for line in out.strip().splitlines():
        f = [x.strip() for x in line.split(",")]
        if len(f) < 9:
            continue
        try:
            ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
            rows.append((ts, float(f[1]), float(f[2]), [n for n, v in zip(names, f[5:9]) if v.lower().startswith("active")]))
        except ValueError:
            continue
    if not rows:
        return None
    inside = [r for r in rows if t_begin <= r[0] <= t_end]
    window = "timed region"
    if not inside:
        inside, window = rows, "warm-up + timed region (timed region shorter than the sampling period)"
    reasons = sorted({x for r in inside for x in r[3]})
    return {"sm_mhz": statistics.median(r[1] for r in inside), "sm_max_mhz": max(r[2] for r in inside), "samples": len(inside),
            "window": window, "reasons": reasons}


def bind_to_gpu_numa(torch, local_rank):
    """Several ranks on one host: run this rank — and so first-touch the pinned staging buffers it allocates — on the cores NVML
    reports as local to its GPU, so that eight ranks' host<->device copies do not all cross the socket interconnect from one
    memory node (SCALE_r01: e2e efficiency 0.82 at 8 GPUs with device-resident efficiency 0.97).  Returns the CPU list, or None
    when nothing was changed (no NVML, one node, a cpuset that does not meet the GPU's cores, H2A_BENCH_NUMA=0)."""
    if os.environ.get("H2A_BENCH_NUMA", "1") == "0":
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local_rank)
        try:
            h = pynvml.nvmlDeviceGetHandleByPciBusId("%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id))
        except Exception:   # noqa: BLE001
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (max(os.cpu_count() or 1, 1) + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if len(cpus) < 2 or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:   # noqa: BLE001 — placement is an optimisation, never a reason to fail
        return None


def oracle_sum(orc, points64):
    """Sum of 64-byte affine points with the oracle's own group law (rank-ordered, like the library's combine)."""
    acc = np.zeros(64, np.uint8)
    for i in range(points64.size // 64):
        acc = orc.g1_add(acc, points64[64 * i:64 * i + 64])
    return acc


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import loader as orc
    threads = orc.hw_threads()
    log_n = args.ref_log_n if args.ref_log_n else args.log_n
    n = 1 << log_n
    bases, scalars = orc.gen_bases(1, n), orc.gen_scalars(2, n)
    for _ in range(args.warmup):
        orc.msm(bases, scalars, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.msm(bases, scalars, threads=threads)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32x8 (254-bit Montgomery)", "data": "synthetic",
            "config": {"workload": "msm_g1 2^%d points" % args.log_n,
                       "note": "CPU restatement (oracle/) of halo2 best_multiexp, not the reference binary: the reference "
                               "cannot be built here (no Rust; un-vendored git dependencies)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "each step = one best_multiexp over the first 2^%d points of the workload" % log_n},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=22)
    ap.add_argument("--ref-log-n", type=int, default=0, help="points per step of the CPU reference arm (0 = --log-n, the same workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU timings (the oracle parity gate still runs)")
    ap.add_argument("--no-precompute", action="store_true", help="plain Pippenger without the per-Params window tables")
    ap.add_argument("--no-prove", action="store_true", help="skip the k=20 prover pipeline measurement (the metric's first half)")
    ap.add_argument("--no-extras", action="store_true", help="skip the strong / ntt / batch / e2e_cold blocks")
    ap.add_argument("--prove-k", type=int, default=20)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import halo2_aggregation_b200 as h2a
    from oracle import loader as orc      # the checker (parity gate) and the cpu_baseline leg only

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GPU arm has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dist = None
    numa_cpus = bind_to_gpu_numa(torch, local_rank) if world > 1 else None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = h2a.Context(local_rank)
    if world > 1:
        ctx.comm_init_torch()     # the library's own NCCL communicators (csrc/comm.cu); torch.distributed only carries the unique ids
    ctx.set_profiling(True)
    n = 1 << args.log_n
    cpu_threads = max(1, orc.hw_threads() // world)     # all ranks check at once

    def combine(partial):
        # 64-byte affine partials allgathered as raw bytes by the library (h2a_comm_allgather) and summed in rank order
        return h2a.g1_sum(ctx.comm_allgather(partial)) if world > 1 else partial

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    stream = torch.cuda.ExternalStream(ctx.stream)

    def timed(fn, steps, kind=0):
        phases = []
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launch_count()
        e0.record(stream)
        t0 = time.perf_counter()
        for _ in range(steps):
            out = fn()
            phases.append(ctx.last_phases(kind))
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        ms = max(e0.elapsed_time(e1), 0.0)
        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), ctx.launch_count() - l0, phases, out

    def all_ranks_ok(ok):
        if world == 1:
            return ok
        t = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(int(t[0]))

    # ---------------------------------------------------------------------------------------------- headline (weak)
    # this rank's point range [rank*n, (rank+1)*n) of the global MSM, generated in place in HBM
    d_bases = torch.empty(64 * n, dtype=torch.uint8, device="cuda")
    d_scal = torch.empty(32 * n, dtype=torch.uint8, device="cuda")
    ctx.gen_bases_dev(1, n, d_bases.data_ptr(), first=rank * n)
    ctx.gen_scalars_dev(2, n, d_scal.data_ptr(), first=rank * n)
    bases = ctx.bases_from_device(d_bases.data_ptr(), n)
    if not args.no_precompute:
        bases.precompute(-1)   # window tables 2^(c*w) * P_i, built once per Params (h2a_bases_precompute)

    def step_dev():
        return combine(ctx.msm_dev(bases, d_scal.data_ptr(), n))

    host_scal = torch.empty(32 * n, dtype=torch.uint8).pin_memory()
    host_scal.copy_(d_scal.cpu())
    host_np = host_scal.numpy()

    def step_e2e():
        return combine(ctx.msm(bases, host_np))

    mon = clocks_monitor_start(local_rank) if rank == 0 else None
    for _ in range(args.warmup):
        step_dev()
    t_begin = time.time()
    ms_total, wall_ms, launches, phases, result = timed(step_dev, args.steps)
    t_end = time.time()
    clocks = clocks_monitor_stop(mon, t_begin, t_end)
    # the same K MSMs issued through the pipelined batch entry point (two lanes), as the prover commits its columns
    def batch_dev():
        return ctx.msm_batch_dev(bases, [d_scal.data_ptr()] * args.steps, [n] * args.steps)
    batch_dev()
    bat_ms, _, _, _, r_bat = timed(batch_dev, 1)
    assert all(bytes(combine(r)) == bytes(result) for r in r_bat), "batched and single results differ"
    for _ in range(2):
        step_e2e()
    e2e_ms, e2e_wall, _, _, r_e2e = timed(step_e2e, args.steps)
    assert bytes(r_e2e) == bytes(result), "e2e and device-resident results differ"

    # ---------------------------------------------------------------------------------------------- parity gate
    # the oracle's best_multiexp over this rank's own inputs (downloaded from HBM: the very bytes the kernels read)
    host_bases = d_bases.cpu().numpy()
    mine = ctx.msm_dev(bases, d_scal.data_ptr(), n)
    t0 = time.perf_counter()
    want_mine = orc.msm(host_bases, host_np, threads=cpu_threads)
    cpu_msm_s = time.perf_counter() - t0
    ok = bytes(mine) == bytes(want_mine)
    if world > 1:
        allw = h2a.allgather_points(want_mine, device="cuda")
        ok = ok and bytes(oracle_sum(orc, allw)) == bytes(result)
    else:
        ok = ok and bytes(result) == bytes(want_mine)
    if not all_ranks_ok(ok):
        raise SystemExit("bench.py: PARITY FAILURE — the timed MSM result differs from the oracle's best_multiexp on the same inputs")
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_baseline = {"value": n / cpu_msm_s / 1e6, "unit": UNIT, "cores": cpu_threads, "kind": "port",
                        "sample": "the whole workload: one best_multiexp (oracle/ restatement) over the same 2^%d points, %.2f s" % (args.log_n, cpu_msm_s)}
    del host_bases

    total_pts = n * world
    ms_per_step = ms_total / args.steps
    value = total_pts / (ms_per_step * 1e-3) / 1e6
    e2e_value = total_pts / (e2e_ms / args.steps * 1e-3) / 1e6

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    imad_peak = ctx.bench_imad()
    modmul_peak = ctx.bench_modmul()

    line = None
    if rank == 0:
        names = [p[0] for p in phases[0]]
        avg = {nm: sum(ph[i][1] for ph in phases) / len(phases) for i, nm in enumerate(names)}
        acc_ms = avg.get("accumulate+merge", 0.0)
        if args.no_precompute:
            c = 16 if args.log_n >= 17 else 15
        else:
            c = 20 if args.log_n >= 16 else 16
        windows = (254 + c - 1) // c
        adds = n * windows
        achieved_gbs = ALGO_BYTES_PER_POINT * n / (acc_ms * 1e-3) / 1e9 if acc_ms else 0.0
        giga_mul = adds * MODMUL_PER_ADD / (acc_ms * 1e-3) / 1e9 if acc_ms else 0.0
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "accumulate_dram_bytes.json")) as f:
                traffic = json.load(f).get("bytes_per_msm_2^%d" % args.log_n)
        except (OSError, ValueError):
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32x8 (254-bit Montgomery)", "data": "synthetic",
            "parity_checked": "oracle",
            "parity": "timed result == oracle best_multiexp over the same 2^%d points per rank (and the oracle's sum of the per-rank results), "
                      "checked before this line was printed" % args.log_n,
            "config": {"workload": "msm_g1 2^%d points per GPU (point-range shard of one %d-point MSM; 64-B partials allgathered and summed)" % (args.log_n, total_pts),
                       "curve": "BN254 G1", "window_bits": c, "windows": windows,
                       "bases": "resident in HBM" + ("" if args.no_precompute else " with per-Params window tables 2^(c*w)*P_i (%.1f GB, built once by h2a_bases_precompute, outside the timed region; see e2e_cold)" % (windows * n * 64 / 1e9)),
                       "l2": "inputs %.0f MB per step exceed the 126 MB L2" % (ALGO_BYTES_PER_POINT * n / 1e6),
                       "e2e_bases": "resident in HBM (uploaded once, like Params); scalars come from pinned host memory every step",
                       "host_placement": ("rank 0 runs on the %d cores NVML reports as local to its GPU (pinned buffers first-touched there); "
                                          "every rank does the same" % len(numa_cpus)) if numa_cpus else "default"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 128 * 64,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches,
            "wall_ms_per_step": wall_ms / args.steps,
            "batched": {"value": total_pts / (bat_ms / args.steps * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": bat_ms / args.steps,
                        "note": "the same %d MSMs through h2a_msm_g1_batch_dev (two pipelined lanes); not the headline value" % args.steps},
            "phases_ms": avg,
            "roofline": {"kernel": "bucket accumulation: batched-affine addition tree (aff_* kernels) + msm_accumulate_pts_kernel",
                         "bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved_gbs / hbm_peak if hbm_peak else None, "traffic": traffic, "peak_source": peak_src,
                         "launch_ms": acc_ms,
                         "note": "algorithmic 96 B/point over the duration of the accumulation kernels; the phase is bound by the "
                                 "integer pipe (254-bit Montgomery products on IMAD.WIDE), see int_pipe"},
            "int_pipe": {"kernel": "bucket accumulation", "achieved": giga_mul, "peak": modmul_peak, "unit": "1e9 Montgomery products/s",
                         "frac": giga_mul / modmul_peak if modmul_peak else None,
                         "peak_source": "h2a_bench_modmul micro-benchmark, same run (IMAD.WIDE-bound; 32-bit IMAD peak %.1f T/s)" % imad_peak,
                         "algorithmic": "%d additions x %d products (batched affine)" % (adds, MODMUL_PER_ADD)},
            "cpu_baseline": cpu_baseline,
        }

    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_extras
    from oracle import plonk as pk        # the oracle's verifier: checker of the proofs the blocks below produce
    from oracle import pymodel as pm
    from oracle import mulvar as mv       # checker of the witness cells of the mul_var block
    env = dict(ctx=ctx, h2a=h2a, orc=orc, pk=pk, pm=pm, mv=mv, torch=torch, dist=dist, rank=rank, world=world, local_rank=local_rank, args=args,
               timed=timed, barrier=barrier, combine=combine, all_ranks_ok=all_ranks_ok, hbm_peak=hbm_peak, peak_src=peak_src,
               modmul_peak=modmul_peak, cpu_threads=cpu_threads)
    extras = {}

    def block(name, fn, *a, **kw):
        """One block beside the headline.  A parity failure (SystemExit) always ends the run without a line.  Any other failure
        of a single-GPU block — out of memory, a full /tmp — is recorded in that block's place and the headline line is still
        printed; with several ranks a failure must stay fatal (the other ranks would wait in the block's next collective)."""
        try:
            extras[name] = fn(*a, **kw)
        except Exception as e:   # noqa: BLE001 — SystemExit (parity) is not an Exception and passes through
            if world > 1:
                raise
            extras[name] = {"error": "%s: %s" % (type(e).__name__, e)}
            print("bench.py: block %r failed: %s: %s" % (name, type(e).__name__, e), file=sys.stderr, flush=True)

    if not args.no_extras:
        # strong scaling reuses the headline's resident inputs at N = 1
        block("strong", bench_extras.strong_block, env, headline=(bases, d_scal, n, value, ms_per_step, result))
        if world == 1:
            block("e2e_cold", bench_extras.cold_block, env, d_bases, host_np, n, result)
    bases.free()
    del d_bases, d_scal, host_scal, host_np
    torch.cuda.empty_cache()
    if not args.no_extras:
        if world == 1:
            block("ntt", bench_extras.ntt_block, env)
            block("sweep", bench_extras.sweep_block, env)
            block("mulvar", bench_extras.mulvar_block, env)
            block("params", bench_extras.params_block, env)
        block("batch", bench_extras.batch_block, env)
    if not args.no_prove:
        block("prove", bench_extras.prove_block, env)
    if line is not None:
        line.update(extras)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
