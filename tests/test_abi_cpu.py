"""CPU-only checks of the C-ABI library: it loads, exports every symbol include/h2agg.h declares,
refuses to run without a GPU (no fallback), and its host-side glue (transcript, point sums, roots of
unity) agrees with the oracle.  No device compute here."""
import os
import random

import numpy as np
import pytest

import halo2_aggregation_b200 as h2a
from oracle import pymodel as pm


def test_library_exports_every_declared_symbol():
    lib = h2a.load_library()
    syms = h2a.declared_symbols()
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.h2a_version() == 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(h2a.H2AError) as e:
        h2a.Context(0)
    assert e.value.code == -3
    assert h2a.load_library().h2a_device_count() == 0


def test_root_of_unity_matches_oracle(orc):
    for k in (0, 1, 9, 20, 23, 28):
        assert bytes(h2a.fr_root_of_unity(k)) == bytes(orc.fr_root_of_unity(k))
    with pytest.raises(h2a.H2AError):
        h2a.fr_root_of_unity(29)


def test_transcript_matches_oracle_and_pymodel(orc):
    rng = random.Random(99)
    t, o, m = h2a.Transcript(), orc.Transcript(), pm.Blake2bTranscript()
    for _ in range(60):
        kind = rng.randrange(3)
        if kind == 0:
            pt = pm.g1_mul(pm.G1, rng.randrange(1, pm.R))
            b = np.frombuffer(pm.affine_bytes(pt), dtype=np.uint8)
            t.common_point(b); o.common_point(b); m.common_point(pt)
        elif kind == 1:
            s = rng.randrange(pm.R)
            b = np.frombuffer(pm.fr_mont_bytes(s), dtype=np.uint8)
            t.common_scalar(b); o.common_scalar(b); m.common_scalar(s)
        else:
            got = t.squeeze_challenge()
            assert bytes(got) == bytes(o.squeeze())
            assert pm.fr_from_mont_bytes(got) == m.squeeze_challenge()
    with pytest.raises(h2a.H2AError):
        t.common_point(np.zeros(64, np.uint8))  # the identity cannot be absorbed


def test_g1_sum_matches_oracle(orc):
    pts = orc.gen_bases(3, 9)
    acc = np.zeros(64, np.uint8)
    for i in range(9):
        acc = orc.g1_add(acc, pts[64 * i:64 * i + 64])
    assert bytes(h2a.g1_sum(pts)) == bytes(acc)
    # P + (-P) + O
    p = pm.g1_mul(pm.G1, 5)
    trio = np.frombuffer(pm.affine_bytes(p) + pm.affine_bytes(pm.g1_neg(p)) + bytes(64), dtype=np.uint8)
    assert bytes(h2a.g1_sum(trio)) == bytes(64)
    dbl = np.frombuffer(pm.affine_bytes(p) * 2, dtype=np.uint8)
    assert pm.affine_from_bytes(h2a.g1_sum(dbl)) == pm.g1_mul(pm.G1, 10)
    assert bytes(h2a.g1_sum(np.zeros(0, np.uint8))) == bytes(64)


def test_xorshift_setup_secret_matches_model():
    """The reference seeds its KZG setup with a fixed XorShift seed (examples/simple-example.rs:584-587)."""
    got = h2a.xorshift_scalar(pm.REFERENCE_SETUP_SEED)
    assert pm.fr_from_mont_bytes(got) == pm.setup_secret_from_seed()
    seed = bytes(range(1, 17))
    rng = pm.XorShiftRng(seed)
    first = rng.next_u32()
    x, y, z, w = (int.from_bytes(seed[4 * i:4 * i + 4], "little") for i in range(4))
    t = (x ^ (x << 11)) & 0xFFFFFFFF
    assert first == (w ^ (w >> 19) ^ (t ^ (t >> 8))) & 0xFFFFFFFF
    assert pm.fr_from_mont_bytes(h2a.xorshift_scalar(bytes(16))) == pm.setup_secret_from_seed(bytes(16))   # all-zero seed is replaced
    assert h2a.load_library().h2a_xorshift_scalar(None, None) == -1


def test_argument_errors_without_a_device():
    import ctypes
    lib = h2a.load_library()
    out = np.zeros(64, np.uint8)
    assert lib.h2a_g1_sum(None, ctypes.c_size_t(3), out.ctypes.data_as(ctypes.c_void_p)) == -1
    assert lib.h2a_transcript_common_point(None, None) == -1
    assert lib.h2a_init(None, 0) == -1
    assert lib.h2a_destroy(None) == -1
    assert lib.h2a_blinds_len(None) == 0 and lib.h2a_proof_len(None) == 0
    assert lib.h2a_msm_set_group(None, 8, 2) == -1 and lib.h2a_msm_set_host_split(None, 2) == -1


def test_permutation_assembly_cycles():
    """Host bookkeeping of key generation (h2a_assembly_copy): the mapping stays a permutation and its cycles are
    exactly the classes of the copy constraints, whatever the order and redundancy of the copies."""
    rng = random.Random(5)
    n_cols, k = 5, 6
    n = 1 << k
    asm = h2a.PermutationAssembly(n_cols, k)
    assert list(asm.mapping()) == list(range(n_cols * n))
    parent = list(range(n_cols * n))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    for _ in range(400):
        a, b = rng.randrange(n_cols * n), rng.randrange(n_cols * n)
        if rng.random() < 0.3:
            b = a                                     # a cell copied to itself changes nothing
        asm.copy(a // n, a % n, b // n, b % n)
        parent[find(a)] = find(b)
    nxt = [int(v) for v in asm.mapping()]
    assert sorted(nxt) == list(range(n_cols * n))
    seen = set()
    for start in range(n_cols * n):
        if start in seen:
            continue
        cyc, c = [], start
        while c not in seen:
            seen.add(c); cyc.append(c); c = nxt[c]
        assert c == start
        cls = find(start)
        assert all(find(x) == cls for x in cyc)
        assert len(cyc) == sum(1 for x in range(n_cols * n) if find(x) == cls)
    with pytest.raises(h2a.H2AError):
        asm.copy(n_cols, 0, 0, 0)
    with pytest.raises(h2a.H2AError):
        asm.copy(0, n, 0, 0)


def test_tree_halves_never_share_a_region():
    """The two halves of the MSM's addition tree run on two unordered streams (csrc/msm.cu): whatever one half writes in any
    round must be disjoint from everything the other half reads or writes in any round, each half must read exactly what it
    wrote the round before, and the last round must leave the points contiguous (half 0 then half 1) in the common array."""
    import ctypes
    lib = h2a.load_library()

    def spans(total, R, half, rnd):
        out = (ctypes.c_int64 * 6)()
        assert lib.h2a_tree_layout(ctypes.c_uint64(total), R, half, rnd, out) == 0
        return tuple(out[:3]), tuple(out[3:])

    def overlap(a, b):
        return a[0] == b[0] and a[0] >= 0 and a[1] < b[1] + b[2] and b[1] < a[1] + a[2]

    for R in range(1, 6):
        unit = 2 << R
        for total in (unit, 3 * unit, 17 * unit, 1000 * unit, ((17791234 + unit - 1) // unit) * unit):
            acc = {0: [], 1: []}                                    # every span a half touches, with (span, is_write)
            for half in (0, 1):
                prev_out = None
                for rnd in range(R):
                    i, o = spans(total, R, half, rnd)
                    assert o[2] == (total // 4) >> rnd
                    if rnd == 0:
                        assert i[0] == -1 and i[1] == half * (total // 2) and i[2] == total // 2
                    else:
                        assert i == prev_out                       # reads exactly what it wrote the round before
                        acc[half].append((i, False))
                    acc[half].append((o, True))
                    prev_out = o
                assert prev_out[0] == 2 and prev_out[1] == half * (total >> (R + 1)) and prev_out[2] == total >> (R + 1)
            for a, a_w in acc[0]:
                for b, b_w in acc[1]:
                    assert not ((a_w or b_w) and overlap(a, b)), (R, total, a, b)
    bad = (ctypes.c_int64 * 6)()
    assert lib.h2a_tree_layout(ctypes.c_uint64(96), 5, 0, 0, bad) == -1    # not a multiple of 2^(R+1)
    assert lib.h2a_tree_layout(ctypes.c_uint64(64), 5, 0, 5, bad) == -1    # round out of range


# ---- the boundary as other languages see it: the Rust crates and a plain C program
def test_vk_hash_matches_hashlib_and_model():
    """Row a9: the scalar the verifier absorbs for the verifying key (src/verifier.rs:341-358) — Blake2b-512 with personal
    "Halo2-Verify-Key" over len_le64 || Debug string, then Fr::from_bytes_wide — from the library, the oracle's model and hashlib
    directly, at lengths on both sides of the 128-byte block boundaries (the length prefix shifts them by 8)."""
    import hashlib
    rng = random.Random(9)
    for ln in (0, 1, 7, 119, 120, 121, 127, 128, 129, 247, 248, 249, 1000, 5000):
        data = bytes(rng.randrange(32, 127) for _ in range(ln))
        h = hashlib.blake2b(ln.to_bytes(8, "little") + data, digest_size=64, person=b"Halo2-Verify-Key").digest()
        want = int.from_bytes(h, "little") % pm.R
        assert pm.vk_hash_from_pinned(data) == want
        assert pm.fr_from_mont_bytes(h2a.vk_hash(data)) == want, ln
    assert pm.fr_from_mont_bytes(h2a.vk_hash("PinnedVerificationKey { k: 9 }")) == pm.vk_hash_from_pinned("PinnedVerificationKey { k: 9 }")
    assert h2a.load_library().h2a_vk_hash(None, 5, None) == -1


def _run_tool(*args):
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return subprocess.run([sys.executable] + [os.path.join(root, a) if a.endswith(".py") else a for a in args], capture_output=True, text=True, cwd=root)


def test_rust_sys_crate_matches_header():
    """h2agg-sys/src/lib.rs is generated from include/h2agg.h: the committed file is current, and it declares exactly the
    functions the header declares (which test_every_declared_symbol_is_exported ties to the built library)."""
    import re
    r = _run_tool("tools/gen_rust_sys.py", "--check")
    assert r.returncode == 0, r.stdout + r.stderr
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "h2agg-sys", "src", "lib.rs")).read()
    in_crate = sorted(set(re.findall(r"pub fn (h2a_[a-z0-9_]+)\(", text)))
    assert in_crate == h2a.declared_symbols()
    for opaque in ("h2a_ctx", "h2a_bases", "h2a_circuit", "h2a_assembly", "h2a_transcript"):
        assert "pub struct %s" % opaque in text
    assert "pub type h2a_exchange_fn" in text and "pub const H2A_ERR_PROOF: c_int = -5;" in text


def test_rust_shim_uses_only_declared_symbols():
    """Every sys::h2a_* the safe wrappers call exists in the header, and the wrappers cover the reference's surface
    (best_multiexp, best_fft, EvaluationDomain, Params read / write / verifier_params, create_proof, verify_proof)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "h2agg-shim", "src", "lib.rs")).read()
    used = set(re.findall(r"sys::(h2a_[a-z0-9_]+)\(", text))
    assert used and used <= set(h2a.declared_symbols()), used - set(h2a.declared_symbols())
    for name in ("pub fn best_multiexp", "pub fn best_fft", "pub struct EvaluationDomain", "pub fn lagrange_to_coeff", "pub fn coeff_to_extended",
                 "pub fn extended_to_coeff", "pub fn setup", "pub fn read", "pub fn write", "pub fn verifier_params", "pub fn commit_lagrange",
                 "pub fn create_proof", "pub fn verify_proof", "pub fn verify_accumulate", "pub struct Transcript"):
        assert name in text, name


def test_header_is_valid_c_and_smoke_program_compiles(tmp_path):
    """include/h2agg.h is C, not just C++: the plain C smoke program compiles and links against the built library (it runs
    on the GPU box, tests/test_gpu_abi_c.py); its known answers are current with the oracle's big-integer model."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = _run_tool("tests/c/gen_abi_smoke_kat.py", "--check")
    assert r.returncode == 0, r.stdout + r.stderr
    exe = str(tmp_path / "abi_smoke")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(root, "include"), os.path.join(root, "tests", "c", "abi_smoke.c"),
                        "-L" + os.path.dirname(h2a.library_path()), "-lh2agg", "-Wl,-rpath," + os.path.dirname(h2a.library_path()), "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def _build_cpp_example(tmp_path):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "simple_example")
    libdir = os.path.dirname(h2a.library_path())
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(root, "include"),
                        os.path.join(root, "examples", "simple_example.cpp"), "-L" + libdir, "-lh2agg", "-Wl,-rpath," + libdir, "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_cpp_host_side_compiles_and_its_host_half_runs(tmp_path):
    """include/h2agg.hpp — the host side above the C ABI in C++, the compiled twin of the Rust shim — and the C++ restatement of
    the reference's example (examples/simple_example.cpp) build warning-free against the library; the half of the example that
    needs no GPU (field helpers, 68-bit limb packing, circuit description, copy constraints -> permutation, transcript, the
    XorShift secret) agrees with the oracle's known answers, which are current.  The GPU half: tests/test_gpu_abi_cpp.py."""
    import subprocess
    r = _run_tool("tests/c/gen_simple_example_kat.py", "--check")
    assert r.returncode == 0, r.stdout + r.stderr
    exe = _build_cpp_example(tmp_path)
    r = subprocess.run([exe, "--host-only"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "host half: all ok" in r.stdout and "FAIL" not in r.stdout, r.stdout + r.stderr


def test_cpp_host_side_covers_the_rust_shim_surface():
    """The C++ header and the Rust shim are two statements of one interface: every h2a_* entry point the shim reaches is reached
    by the header too, and both carry the dependency's names."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hpp = open(os.path.join(root, "include", "h2agg.hpp")).read()
    shim = open(os.path.join(root, "h2agg-shim", "src", "lib.rs")).read()
    used_hpp = set(re.findall(r"\b(h2a_[a-z0-9_]+)\(", hpp))
    used_shim = set(re.findall(r"sys::(h2a_[a-z0-9_]+)\(", shim))
    assert used_hpp <= set(h2a.declared_symbols()), used_hpp - set(h2a.declared_symbols())
    assert used_shim <= used_hpp, used_shim - used_hpp
    for name in ("best_multiexp", "best_fft", "EvaluationDomain", "lagrange_to_coeff", "coeff_to_extended", "extended_to_coeff", "get_omega",
                 "get_quotient_poly_degree", "verifier_params", "commit_lagrange", "create_proof", "verify_proof", "verify_proof_batch",
                 "verify_accumulate", "fold_h", "Transcript", "mul_var_witness", "comm_init", "allgather"):
        assert name in hpp and name in shim, name


def test_generated_field_square_is_current_and_checked():
    """csrc/field_mul_gen.cuh is what tools/gen_field_mul.py emits, and the generator's op lists still pass their check
    against Python big integers (random and edge operands, no carry dropped)."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_field_mul", os.path.join(root, "tools", "gen_field_mul.py"))
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    assert g.check(n_random=300)
    parts = [g.HEADER]
    for field, p_int in (("0", g.FQ), ("1", g.FR)):
        parts.append(g.emit(g.build("sqr", p_int), "mont_sqr_wide", "sqr", field))
    parts.append("}  // namespace h2a\n")
    assert open(g.OUT).read() == "\n\n".join(parts)


def test_distributed_prover_windows_cover_what_a_rank_reads():
    """One proof over several GPUs: rank r evaluates quotient rows [r m/W, (r+1) m/W) and reads its columns at row + rot * step, so
    the window of an extended column it is sent (h2a_dist_window, csrc/dist_layout.hpp) must hold every such row; the pieces are
    contiguous, inside the array, disjoint, and together exactly slice + 2 halo rows."""
    import ctypes
    lib = h2a.load_library()
    out = (ctypes.c_uint32 * 4)()
    for m in (1 << 12, 1 << 16, 3 << 10, 1 << 22):
        for world in (2, 3, 4, 8):
            if m % world:
                continue
            slice_ = m // world
            for halo in (0, 4, 24, slice_ // 4, slice_ // 2):
                held = []
                for rank in range(world):
                    n = lib.h2a_dist_window(ctypes.c_uint32(m), world, rank, ctypes.c_uint32(halo), out)
                    assert n in (1, 2)
                    pieces = [(out[0], out[1]), (out[2], out[3])][:n]
                    rows = set()
                    for first, count in pieces:
                        assert count > 0 and first + count <= m
                        new = set(range(first, first + count)) if m <= (1 << 16) else None
                        if new is not None:
                            assert not (rows & new)
                            rows |= new
                    assert sum(c for _, c in pieces) == slice_ + 2 * halo
                    if m <= (1 << 16):
                        lo, hi = rank * slice_, (rank + 1) * slice_
                        for i in (lo, lo + 1, (lo + hi) // 2, hi - 1):
                            for d in (-halo, -min(1, halo), 0, min(1, halo), halo):
                                assert (i + d) % m in rows, (m, world, rank, halo, i, d)
                        assert rows == {x % m for x in range(lo - halo, hi + halo)}
                    held.append(pieces)
    bad = (ctypes.c_uint32 * 4)()
    assert lib.h2a_dist_window(ctypes.c_uint32(1 << 12), 1, 0, ctypes.c_uint32(4), bad) == -1        # one rank needs no window
    assert lib.h2a_dist_window(ctypes.c_uint32(1 << 12), 3, 0, ctypes.c_uint32(4), bad) == -1        # m not a multiple of world
    assert lib.h2a_dist_window(ctypes.c_uint32(1 << 12), 4, 0, ctypes.c_uint32(600), bad) == -1      # reach wider than half a slice
