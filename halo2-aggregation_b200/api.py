"""ctypes binding of libh2agg.so and a host-side mirror of the reference-facing interface.

Reference interface mirrored (all [UPSTREAM-INFERRED] signatures of the `halo2`
dependency pinned at Cargo.toml:12):
  best_multiexp(coeffs, bases) -> point          call sites: examples/simple-example.rs:638-640
  best_fft(a, omega, log_n)                       domain touched at src/verifier.rs:252,431
  EvaluationDomain::{lagrange_to_coeff, coeff_to_extended, extended_to_coeff, get_omega}
  MultiopenChip::calc_witness native values       src/multiopen.rs:271-509 -> verify_accumulate
  TranscriptChip / Blake2bWrite                   src/transcript.rs:66-129 -> Transcript
Buffers are numpy uint8 arrays in the library's wire layout (32-byte Montgomery field elements,
64-byte affine points).
"""
import ctypes
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_u8p = ctypes.c_void_p
c_sz = ctypes.c_size_t


class H2AError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("h2agg error %d: %s" % (code, msg))
        self.code = code


def library_path():
    """The in-tree build; H2A_LIB names another build of the SAME library (kernel A/B variants made by tools/build_variant.py)."""
    return os.environ.get("H2A_LIB") or os.path.join(_HERE, "libh2agg.so")


def header_path():
    return os.path.join(os.path.dirname(_HERE), "include", "h2agg.h")


def declared_symbols():
    """Every function include/h2agg.h declares."""
    with open(header_path()) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(h2a_[a-z0-9_]+)\s*\(", text)))


def load_library():
    """Loads libh2agg.so; fails loudly when it has not been built (no fallback of any kind)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise H2AError(-3, "libh2agg.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = ctypes.CDLL(path)
    lib.h2a_last_error.restype = ctypes.c_char_p
    lib.h2a_stream.restype = ctypes.c_void_p
    lib.h2a_phase_name.restype = ctypes.c_char_p
    lib.h2a_launch_count.restype = ctypes.c_uint64
    lib.h2a_bases_len.restype = c_sz
    lib.h2a_blinds_len.restype = c_sz
    lib.h2a_mulvar_witness_len.restype = c_sz
    lib.h2a_proof_len.restype = c_sz
    lib.h2a_prove_phase_name.restype = ctypes.c_char_p
    lib.h2a_transcript_new.restype = ctypes.c_void_p
    lib.h2a_transcript_free.restype = None
    lib.h2a_assembly_free.restype = None
    lib.h2a_assembly_free.argtypes = [ctypes.c_void_p]
    _LIB = lib
    return lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return ctypes.c_void_p(a)
    return a.ctypes.data_as(ctypes.c_void_p)


def _bytes(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def fr_root_of_unity(k):
    out = np.zeros(32, np.uint8)
    rc = load_library().h2a_fr_root_of_unity(ctypes.c_uint32(k), _ptr(out))
    if rc != 0:
        raise H2AError(rc, "root_of_unity(%d)" % k)
    return out


def vk_hash(pinned_debug):
    """h2a_vk_hash: the transcript scalar of a verifying key from `format!("{:?}", vk.pinned())` (src/verifier.rs:341-358)."""
    data = np.frombuffer(pinned_debug.encode() if isinstance(pinned_debug, str) else bytes(pinned_debug), dtype=np.uint8)
    out = np.zeros(32, np.uint8)
    rc = load_library().h2a_vk_hash(_ptr(data) if data.size else None, c_sz(data.size), _ptr(out))
    if rc != 0:
        raise H2AError(rc, "vk_hash")
    return out


def xorshift_scalar(seed16):
    """The KZG secret the reference's seeded XorShiftRng yields (examples/simple-example.rs:584-589)."""
    seed = np.frombuffer(bytes(seed16), dtype=np.uint8)
    out = np.zeros(32, np.uint8)
    rc = load_library().h2a_xorshift_scalar(_ptr(seed), _ptr(out))
    if rc != 0:
        raise H2AError(rc, "xorshift_scalar")
    return out


def comm_unique_id():
    """128 opaque bytes naming a new communicator (rank 0 draws two and hands them to every rank)."""
    out = np.zeros(128, np.uint8)
    rc = load_library().h2a_comm_unique_id(_ptr(out))
    if rc != 0:
        raise H2AError(rc, "h2a_comm_unique_id failed (libnccl.so.2 not found?)")
    return out


def g1_sum(points):
    pts = _bytes(points)
    out = np.zeros(64, np.uint8)
    rc = load_library().h2a_g1_sum(_ptr(pts), c_sz(pts.size // 64), _ptr(out))
    if rc != 0:
        raise H2AError(rc, "g1_sum")
    return out


class Context:
    """One per GPU / process (h2a_init)."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = ctypes.c_void_p()
        rc = self.lib.h2a_init(ctypes.byref(h), int(device))
        if rc != 0:
            raise H2AError(rc, "h2a_init(device=%d) failed: no usable sm_100 CUDA device (there is no CPU fallback)" % device)
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.h2a_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise H2AError(rc, (self.lib.h2a_last_error(self.h) or b"").decode())

    # ---- plumbing
    @property
    def stream(self):
        return self.lib.h2a_stream(self.h)

    def sync(self):
        self._check(self.lib.h2a_sync(self.h))

    def launch_count(self):
        return int(self.lib.h2a_launch_count(self.h))

    def dev_alloc(self, nbytes):
        p = ctypes.c_void_p()
        self._check(self.lib.h2a_dev_alloc(self.h, c_sz(nbytes), ctypes.byref(p)))
        return p.value

    def dev_free(self, p):
        self._check(self.lib.h2a_dev_free(self.h, ctypes.c_void_p(p)))

    def h2d(self, dev, host):
        host = _bytes(host)
        self._check(self.lib.h2a_copy_h2d(self.h, ctypes.c_void_p(dev), _ptr(host), c_sz(host.size)))

    def d2h(self, dev, nbytes):
        out = np.zeros(nbytes, np.uint8)
        self._check(self.lib.h2a_copy_d2h(self.h, _ptr(out), ctypes.c_void_p(dev), c_sz(nbytes)))
        return out

    def set_profiling(self, on):
        self._check(self.lib.h2a_set_profiling(self.h, int(bool(on))))

    def last_phases(self, kind=0):
        buf = (ctypes.c_float * 16)()
        n = self.lib.h2a_last_phase_ms(self.h, buf, 16)
        return [((self.lib.h2a_phase_name(kind, i) or b"").decode(), float(buf[i])) for i in range(max(n, 0))]

    def bench_imad(self):
        v = ctypes.c_double()
        self._check(self.lib.h2a_bench_imad(self.h, ctypes.byref(v)))
        return v.value

    def bench_modmul(self):
        v = ctypes.c_double()
        self._check(self.lib.h2a_bench_modmul(self.h, ctypes.byref(v)))
        return v.value

    # ---- synthetic inputs
    def gen_scalars_dev(self, seed, n, dev, first=0):
        self._check(self.lib.h2a_gen_scalars_dev(self.h, ctypes.c_uint64(seed), c_sz(first), c_sz(n), ctypes.c_void_p(dev)))

    def gen_bases_dev(self, seed, n, dev, first=0):
        self._check(self.lib.h2a_gen_bases_dev(self.h, ctypes.c_uint64(seed), c_sz(first), c_sz(n), ctypes.c_void_p(dev)))

    # ---- MSM
    def upload_bases(self, affine):
        affine = _bytes(affine)
        h = ctypes.c_void_p()
        self._check(self.lib.h2a_bases_upload(self.h, _ptr(affine), c_sz(affine.size // 64), ctypes.byref(h)))
        return Bases(self, h)

    def bases_from_device(self, dev, n):
        h = ctypes.c_void_p()
        self._check(self.lib.h2a_bases_from_device(self.h, ctypes.c_void_p(dev), c_sz(n), ctypes.byref(h)))
        return Bases(self, h)

    def kzg_setup(self, k, s):
        """Setup::<Bn256>::new(k, .) for the secret s: (g, g_lagrange) resident handles."""
        g, gl = ctypes.c_void_p(), ctypes.c_void_p()
        self._check(self.lib.h2a_kzg_setup(self.h, ctypes.c_uint32(k), _ptr(_bytes(s)), ctypes.byref(g), ctypes.byref(gl)))
        return Bases(self, g), Bases(self, gl)

    def params_write(self, path, k, g, g_lagrange, compressed=False, trailer=None):
        """Params::write: stream (g, g_lagrange) to a parameter file (format: csrc/params.cu)."""
        tr = None if trailer is None else _bytes(trailer)
        if tr is not None and tr.size != 128:
            raise ValueError("trailer must be 128 bytes")
        self._check(self.lib.h2a_params_write(self.h, os.fsencode(path), ctypes.c_uint32(k), g.h, g_lagrange.h, int(bool(compressed)), _ptr(tr)))

    def params_read(self, path):
        """Params::read: (k, g, g_lagrange, trailer or None); digest and curve membership are checked on load."""
        k, g, gl, has = ctypes.c_uint32(), ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int()
        tr = np.zeros(128, np.uint8)
        self._check(self.lib.h2a_params_read(self.h, os.fsencode(path), ctypes.byref(k), ctypes.byref(g), ctypes.byref(gl), _ptr(tr), ctypes.byref(has)))
        return int(k.value), Bases(self, g), Bases(self, gl), (tr if has.value else None)

    def verifier_params(self, g_lagrange, public_inputs_size):
        """Setup::verifier_params: a view of the first `public_inputs_size` Lagrange bases (shares g_lagrange's memory)."""
        h = ctypes.c_void_p()
        self._check(self.lib.h2a_params_verifier_view(self.h, g_lagrange.h, c_sz(public_inputs_size), ctypes.byref(h)))
        return Bases(self, h)

    # ---- aggregation-circuit witness generation (row f4)
    def mulvar_witness_len(self):
        return int(self.lib.h2a_mulvar_witness_len())

    def mulvar_witness(self, points, scalars, aux, want_witness=True):
        """Witness cells of m non-native mul_var (csrc/mulvar.cu): (results m*64, witness m*len*32 or None, status u32[m]).
        Raises H2AError when an entry cannot be witnessed; `e.status` then holds the per-entry status."""
        points, scalars, aux = _bytes(points), _bytes(scalars), _bytes(aux)
        m = points.size // 64
        if scalars.size != 32 * m or aux.size != 64:
            raise ValueError("mulvar_witness: m points of 64 bytes, m scalars of 32 bytes, one auxiliary point")
        res = np.zeros(64 * m, np.uint8)
        wit = np.zeros(32 * self.mulvar_witness_len() * m, np.uint8) if want_witness else None
        status = np.zeros(m, np.uint32)
        rc = self.lib.h2a_mulvar_witness(self.h, _ptr(points), _ptr(scalars), c_sz(m), _ptr(aux), _ptr(res), _ptr(wit), _ptr(status))
        if rc != 0:
            err = H2AError(rc, self.lib.h2a_last_error(self.h).decode())
            err.status = status
            raise err
        return res, wit, status

    def mulvar_witness_dev(self, d_points, d_scalars, m, aux, d_results, d_witness):
        status = np.zeros(m, np.uint32)
        self._check(self.lib.h2a_mulvar_witness_dev(self.h, ctypes.c_void_p(d_points), ctypes.c_void_p(d_scalars), c_sz(m), _ptr(_bytes(aux)),
                                                   ctypes.c_void_p(d_results), ctypes.c_void_p(d_witness), _ptr(status)))
        return status

    # ---- several GPUs (the library's own NCCL plumbing, csrc/comm.cu)
    def comm_init(self, rank, world, ids):
        """ids: 256 bytes = two h2a_comm_unique_id() results drawn by rank 0 and handed to every rank."""
        ids = _bytes(ids)
        if ids.size != 256:
            raise ValueError("ids must be two 128-byte unique ids")
        self._check(self.lib.h2a_comm_init(self.h, int(rank), int(world), _ptr(ids[:128].copy()), _ptr(ids[128:].copy())))

    def comm_init_torch(self):
        """comm_init with the ids passed around through an initialised torch.distributed group."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        ids = [np.concatenate([comm_unique_id(), comm_unique_id()]).tobytes() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        self.comm_init(rank, world, np.frombuffer(ids[0], dtype=np.uint8))
        return rank, world

    def comm_traffic(self):
        """(bytes sent, bytes received) by this rank through the library's collectives so far."""
        out = (ctypes.c_uint64 * 2)()
        self._check(self.lib.h2a_comm_traffic(self.h, out))
        return int(out[0]), int(out[1])

    def comm_destroy(self):
        self._check(self.lib.h2a_comm_destroy(self.h))

    def comm_allgather(self, send):
        send = _bytes(send)
        world = int(self.lib.h2a_comm_world(self.h))
        out = np.zeros(send.size * world, np.uint8)
        self._check(self.lib.h2a_comm_allgather(self.h, _ptr(send), _ptr(out), c_sz(send.size)))
        return out

    def comm_allgather_dev(self, d_send, d_recv, nbytes):
        self._check(self.lib.h2a_comm_allgather_dev(self.h, ctypes.c_void_p(d_send), ctypes.c_void_p(d_recv), c_sz(nbytes)))

    def comm_broadcast_dev(self, d_buf, nbytes, root):
        self._check(self.lib.h2a_comm_broadcast_dev(self.h, ctypes.c_void_p(d_buf), c_sz(nbytes), int(root)))

    def msm_sharded(self, local_bases, d_local_scalars, n_local, offset=0):
        """One MSM whose point ranges live on the ranks of the communicator; every rank gets the same 64 bytes."""
        out = np.zeros(64, np.uint8)
        self._check(self.lib.h2a_msm_g1_sharded(self.h, local_bases.h, c_sz(offset), ctypes.c_void_p(d_local_scalars), c_sz(n_local), _ptr(out)))
        return out

    def set_msm_window(self, c):
        self._check(self.lib.h2a_msm_set_window(self.h, int(c)))

    def set_msm_algorithm(self, algo):
        self._check(self.lib.h2a_msm_set_algorithm(self.h, int(algo)))

    def set_msm_host_split(self, pieces):
        self._check(self.lib.h2a_msm_set_host_split(self.h, int(pieces)))

    def set_msm_group(self, cols, cols_host=2):
        self._check(self.lib.h2a_msm_set_group(self.h, int(cols), int(cols_host)))

    def msm(self, bases, scalars, offset=0):
        """best_multiexp over resident bases, host scalars.  Returns the 64-byte affine result."""
        scalars = _bytes(scalars)
        out = np.zeros(64, np.uint8)
        self._check(self.lib.h2a_msm_g1(self.h, bases.h, c_sz(offset), _ptr(scalars), c_sz(scalars.size // 32), _ptr(out)))
        return out

    def msm_dev(self, bases, d_scalars, n, offset=0):
        out = np.zeros(64, np.uint8)
        self._check(self.lib.h2a_msm_g1_dev(self.h, bases.h, c_sz(offset), ctypes.c_void_p(d_scalars), c_sz(n), _ptr(out)))
        return out

    def msm_batch(self, bases, columns):
        cols = [_bytes(c) for c in columns]
        m = len(cols)
        ptrs = (ctypes.c_void_p * m)(*[c.ctypes.data for c in cols])
        ns = (c_sz * m)(*[c.size // 32 for c in cols])
        out = np.zeros(64 * m, np.uint8)
        self._check(self.lib.h2a_msm_g1_batch(self.h, bases.h, ptrs, ns, m, _ptr(out)))
        return out.reshape(m, 64)

    def msm_batch_dev(self, bases, d_columns, ns):
        m = len(d_columns)
        ptrs = (ctypes.c_void_p * m)(*d_columns)
        nn = (c_sz * m)(*ns)
        out = np.zeros(64 * m, np.uint8)
        self._check(self.lib.h2a_msm_g1_batch_dev(self.h, bases.h, ptrs, nn, m, _ptr(out)))
        return out.reshape(m, 64)

    def msm_adhoc(self, bases_affine, scalars):
        b, s = _bytes(bases_affine), _bytes(scalars)
        if b.size // 64 != s.size // 32:
            raise H2AError(-1, "msm_adhoc: %d bases vs %d scalars" % (b.size // 64, s.size // 32))
        out = np.zeros(64, np.uint8)
        self._check(self.lib.h2a_msm_g1_adhoc(self.h, _ptr(b), _ptr(s), c_sz(s.size // 32), _ptr(out)))
        return out

    # ---- NTT
    def ntt(self, a, log_n, omega, inverse=False, coset_shift=None, inplace=False):
        """`inplace`: transform the caller's buffer (e.g. pinned memory) as the C ABI does, instead of a copy."""
        a = _bytes(a) if inplace else _bytes(a).copy()
        if a.size != 32 << log_n:
            raise H2AError(-1, "ntt: buffer has %d bytes, expected %d" % (a.size, 32 << log_n))
        cs = _bytes(coset_shift) if coset_shift is not None else None
        self._check(self.lib.h2a_ntt(self.h, _ptr(a), ctypes.c_uint32(log_n), _ptr(_bytes(omega)), int(bool(inverse)), _ptr(cs)))
        return a

    def ntt_dev(self, d_a, log_n, omega, inverse=False, coset_shift=None):
        cs = _bytes(coset_shift) if coset_shift is not None else None
        self._check(self.lib.h2a_ntt_dev(self.h, ctypes.c_void_p(d_a), ctypes.c_uint32(log_n), _ptr(_bytes(omega)),
                                         int(bool(inverse)), _ptr(cs)))

    def coeff_to_extended(self, coeffs, k, ext_k, coset_shift, out=None):
        """`out`: optional caller buffer of 32 << ext_k bytes (e.g. pinned memory), as the C ABI takes it."""
        coeffs = _bytes(coeffs)
        if coeffs.size != 32 << k:
            raise H2AError(-1, "coeff_to_extended: buffer has %d bytes, expected %d" % (coeffs.size, 32 << k))
        if out is None:
            out = np.empty(32 << ext_k, np.uint8)
        elif out.dtype != np.uint8 or out.size != 32 << ext_k or not out.flags["C_CONTIGUOUS"]:
            raise H2AError(-1, "coeff_to_extended: out must be a contiguous uint8 buffer of %d bytes" % (32 << ext_k))
        self._check(self.lib.h2a_coeff_to_extended(self.h, _ptr(coeffs), ctypes.c_uint32(k), ctypes.c_uint32(ext_k),
                                                   _ptr(_bytes(coset_shift)), _ptr(out)))
        return out

    def coeff_to_extended_dev(self, d_coeffs, k, ext_k, coset_shift, d_out):
        self._check(self.lib.h2a_coeff_to_extended_dev(self.h, ctypes.c_void_p(d_coeffs), ctypes.c_uint32(k), ctypes.c_uint32(ext_k),
                                                       _ptr(_bytes(coset_shift)), ctypes.c_void_p(d_out)))

    def extended_to_coeff_dev(self, d_ext, ext_k, coset_shift):
        self._check(self.lib.h2a_extended_to_coeff_dev(self.h, ctypes.c_void_p(d_ext), ctypes.c_uint32(ext_k), _ptr(_bytes(coset_shift))))

    def extended_to_coeff(self, ext, ext_k, coset_shift):
        ext = _bytes(ext).copy()
        self._check(self.lib.h2a_extended_to_coeff(self.h, _ptr(ext), ctypes.c_uint32(ext_k), _ptr(_bytes(coset_shift))))
        return ext

    # ---- verifier glue
    def verify_accumulate(self, commitments, rotations, evals, ws, x, u, v, omega, g1):
        c, e, w = _bytes(commitments), _bytes(evals), _bytes(ws)
        rot = np.ascontiguousarray(rotations, dtype=np.int32)
        out = np.zeros(256, np.uint8)
        self._check(self.lib.h2a_verify_accumulate(self.h, _ptr(c), _ptr(rot), _ptr(e), c_sz(rot.size), _ptr(w),
                                                   c_sz(w.size // 64), _ptr(_bytes(x)), _ptr(_bytes(u)), _ptr(_bytes(v)),
                                                   _ptr(_bytes(omega)), _ptr(_bytes(g1)), _ptr(out)))
        return out

    def verify_accumulate_batch(self, proofs, omega, g1):
        """proofs: list of dicts with commitments, rotations, evals, ws, x, u, v."""
        n = len(proofs)
        c = np.concatenate([_bytes(p["commitments"]) for p in proofs]) if n else np.zeros(0, np.uint8)
        r = np.concatenate([np.asarray(p["rotations"], dtype=np.int32) for p in proofs]) if n else np.zeros(0, np.int32)
        e = np.concatenate([_bytes(p["evals"]) for p in proofs]) if n else np.zeros(0, np.uint8)
        w = np.concatenate([_bytes(p["ws"]) for p in proofs]) if n else np.zeros(0, np.uint8)
        xuv = np.concatenate([np.concatenate([_bytes(p["x"]), _bytes(p["u"]), _bytes(p["v"])]) for p in proofs]) if n else np.zeros(0, np.uint8)
        q_off = np.zeros(n + 1, dtype=np.uint64)
        w_off = np.zeros(n + 1, dtype=np.uint64)
        for i, p in enumerate(proofs):
            q_off[i + 1] = q_off[i] + len(p["rotations"])
            w_off[i + 1] = w_off[i] + _bytes(p["ws"]).size // 64
        out = np.zeros(256 * n, np.uint8)
        self._check(self.lib.h2a_verify_accumulate_batch(self.h, c_sz(n), _ptr(c), _ptr(np.ascontiguousarray(r)), _ptr(e), _ptr(q_off),
                                                         _ptr(w), _ptr(w_off), _ptr(xuv), _ptr(_bytes(omega)), _ptr(_bytes(g1)),
                                                         _ptr(out)))
        return out.reshape(n, 256)

    def fold_h(self, h_pieces, xn):
        h = _bytes(h_pieces)
        out = np.zeros(64, np.uint8)
        self._check(self.lib.h2a_fold_h(self.h, _ptr(h), c_sz(h.size // 64), _ptr(_bytes(xn)), _ptr(out)))
        return out

    # ---- test hooks
    FIELD_OPS = {"add": 0, "sub": 1, "mul": 2, "sqr": 3, "inv": 4, "neg": 5, "to_mont": 6, "from_mont": 7, "inv_fast": 8}

    def field_op(self, field, op, a, b=None):
        a = _bytes(a)
        bb = _bytes(b) if b is not None else None
        out = np.zeros(a.size, np.uint8)
        self._check(self.lib.h2a_field_op(self.h, int(field), self.FIELD_OPS[op], _ptr(a), _ptr(bb), _ptr(out), c_sz(a.size // 32)))
        return out

    def g1_op(self, op, a, b=None):
        a = _bytes(a)
        bb = _bytes(b) if b is not None else None
        out = np.zeros(a.size, np.uint8)
        self._check(self.lib.h2a_g1_op(self.h, {"add": 0, "dbl": 1, "dbl_add": 2}[op], _ptr(a), _ptr(bb), _ptr(out), c_sz(a.size // 64)))
        return out


SHAPE_MAGIC = 0x48324153


def serialize_shape(shape):
    """Word stream of csrc/plonk_shape.hpp from any object with the attributes of the reference's vk-derived
    parameters (k, bf, degree, num_*, *_queries, gates, constants, lookups, perm_columns)."""
    w = [SHAPE_MAGIC, shape.k, shape.bf, shape.degree, shape.num_instance, shape.num_advice, shape.num_fixed]
    for qs in (shape.advice_queries, shape.fixed_queries, shape.instance_queries):
        w.append(len(qs))
        for col, rot in qs:
            w += [col, rot & 0xFFFFFFFF]

    def prog(p):
        out = [len(p)]
        for op, arg in p:
            out += [op, arg]
        return out

    w.append(len(shape.gates))
    for g in shape.gates:
        w += prog(g)
    w.append(len(shape.constants))
    w.append(len(shape.lookups))
    for inputs, tables in shape.lookups:
        w.append(len(inputs))
        for p in inputs:
            w += prog(p)
        w.append(len(tables))
        for p in tables:
            w += prog(p)
    w.append(len(shape.perm_columns))
    for t, c, q in shape.perm_columns:
        w += [t, c, q]
    return np.array(w, dtype=np.uint32)


class Circuit:
    """h2a_circuit: a circuit shape + verifying key (+ proving key for `prove`)."""

    def __init__(self, ctx, shape, constants_mont):
        self.ctx = ctx
        words = serialize_shape(shape)
        consts = _bytes(constants_mont) if len(shape.constants) else np.zeros(0, np.uint8)
        h = ctypes.c_void_p()
        ctx._check(ctx.lib.h2a_circuit_create(ctx.h, _ptr(words), c_sz(words.size), _ptr(consts) if consts.size else None,
                                              c_sz(len(shape.constants)), ctypes.byref(h)))
        self.h = h
        self.n_instance, self.n_advice, self.n = shape.num_instance, shape.num_advice, 1 << shape.k

    def free(self):
        if self.h:
            self.ctx._check(self.ctx.lib.h2a_circuit_free(self.ctx.h, self.h))
            self.h = None

    def set_vk(self, fixed_commitments, sigma_commitments, vk_hash):
        f, s = _bytes(fixed_commitments), _bytes(sigma_commitments)
        self.ctx._check(self.ctx.lib.h2a_circuit_set_vk(self.ctx.h, self.h, _ptr(f), _ptr(s), _ptr(_bytes(vk_hash))))

    def set_keys(self, g, g_lagrange, fixed_values, sigmas, vk_hash, coset_shift):
        """Proving key: Params bases handles + fixed / permutation columns (column-major, n elements each)."""
        f, s = _bytes(fixed_values), _bytes(sigmas)
        self.ctx._check(self.ctx.lib.h2a_circuit_set_keys(self.ctx.h, self.h, g.h, g_lagrange.h, _ptr(f) if f.size else None,
                                                          _ptr(s) if s.size else None, _ptr(_bytes(vk_hash)), _ptr(_bytes(coset_shift))))
        self._keep = (g, g_lagrange)

    def set_distribution(self, rank, world, group=None, device=None, native=False):
        """One proof over `world` processes: every rank proves the same inputs.  native=True uses the library's own
        communicator (Context.comm_init): commitments column-parallel, transforms column-parallel with the results
        broadcast, quotient row-parallel.  Otherwise only the commitments are shared and their 64-byte results are
        allgathered through torch.distributed."""
        if world <= 1:
            self.ctx._check(self.ctx.lib.h2a_circuit_set_distribution(self.ctx.h, self.h, 0, 1, None, None))
            self._exchange = None
            return
        if native:
            self.ctx._check(self.ctx.lib.h2a_circuit_set_distribution(self.ctx.h, self.h, int(rank), int(world), None, None))
            self._exchange = None
            return
        from .dist import make_commitment_exchange
        do_exchange = make_commitment_exchange(world, group, device)

        def exchange(_user, buf, m):
            try:
                do_exchange(np.ctypeslib.as_array(ctypes.cast(buf, ctypes.POINTER(ctypes.c_uint8)), shape=(64 * m,)))
                return 0
            except Exception:
                return -1

        cb = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, c_sz)(exchange)
        self._exchange = cb   # keep the callback alive
        self.ctx._check(self.ctx.lib.h2a_circuit_set_distribution(self.ctx.h, self.h, int(rank), int(world), cb, None))

    def get_vk(self, n_fixed, n_perm):
        f, s = np.zeros(64 * n_fixed, np.uint8), np.zeros(64 * n_perm, np.uint8)
        self.ctx._check(self.ctx.lib.h2a_circuit_get_vk(self.ctx.h, self.h, _ptr(f), _ptr(s)))
        return f, s

    def blinds_len(self):
        return int(self.ctx.lib.h2a_blinds_len(self.h))

    def proof_len(self):
        return int(self.ctx.lib.h2a_proof_len(self.h))

    def prove(self, instance_cols, advice_cols, blinds):
        """create_proof: returns (proof bytes, instance commitments)."""
        ic, ac, bl = _bytes(instance_cols), _bytes(advice_cols), _bytes(blinds)
        if bl.size != 32 * self.blinds_len():
            raise H2AError(-1, "prove: blinds has %d elements, expected %d" % (bl.size // 32, self.blinds_len()))
        # the C entry point reads n_instance / n_advice whole columns of 2^k elements
        if ic.size != 32 * self.n * self.n_instance or ac.size != 32 * self.n * self.n_advice:
            raise H2AError(-1, "prove: instance / advice buffers hold %d / %d bytes, expected %d / %d"
                           % (ic.size, ac.size, 32 * self.n * self.n_instance, 32 * self.n * self.n_advice))
        out = np.zeros(self.proof_len(), np.uint8)
        ln = c_sz(0)
        inst = np.zeros(64 * self.n_instance, np.uint8)
        self.ctx._check(self.ctx.lib.h2a_create_proof(self.ctx.h, self.h, _ptr(ic) if ic.size else None, _ptr(ac) if ac.size else None,
                                                      _ptr(bl), _ptr(out), c_sz(out.size), ctypes.byref(ln), _ptr(inst)))
        return bytes(out[:ln.value]), inst

    def prove_phases(self):
        buf = (ctypes.c_float * 32)()
        k = self.ctx.lib.h2a_prove_phase_ms(self.ctx.h, self.h, buf, 32)
        return [((self.ctx.lib.h2a_prove_phase_name(self.ctx.h, i) or b"").decode(), float(buf[i])) for i in range(max(k, 0))]

    def verify(self, instance_commitments, proof):
        """(e, f, w, zw) of one proof: 256 bytes."""
        ic = _bytes(instance_commitments)
        pr = np.frombuffer(bytes(proof), dtype=np.uint8)
        out = np.zeros(256, np.uint8)
        self.ctx._check(self.ctx.lib.h2a_verify_proof(self.ctx.h, self.h, _ptr(ic), _ptr(pr), c_sz(pr.size), _ptr(out)))
        return out

    def verify_batch(self, instance_commitments, proofs):
        ic = _bytes(instance_commitments)
        bufs = [np.frombuffer(bytes(p), dtype=np.uint8) for p in proofs]
        m = len(bufs)
        ptrs = (ctypes.c_void_p * m)(*[b.ctypes.data for b in bufs])
        lens = (c_sz * m)(*[b.size for b in bufs])
        out = np.zeros(256 * m, np.uint8)
        self.ctx._check(self.ctx.lib.h2a_verify_proof_batch(self.ctx.h, self.h, c_sz(m), _ptr(ic), ptrs, lens, _ptr(out)))
        return out.reshape(m, 256)


class Bases:
    """Device-resident affine bases (`Params.g` / `Params.g_lagrange`)."""

    def __init__(self, ctx, h):
        self.ctx, self.h = ctx, h

    def __len__(self):
        return int(self.ctx.lib.h2a_bases_len(self.h))

    def download(self):
        out = np.zeros(64 * len(self), np.uint8)
        self.ctx._check(self.ctx.lib.h2a_bases_download(self.ctx.h, self.h, _ptr(out)))
        return out

    def precompute(self, window_bits=-1):
        """Build the window tables 2^(window_bits*w) * P_i (h2a_bases_precompute); 0 drops them."""
        self.ctx._check(self.ctx.lib.h2a_bases_precompute(self.ctx.h, self.h, int(window_bits)))
        return self

    def free(self):
        if self.h:
            self.ctx._check(self.ctx.lib.h2a_bases_free(self.ctx.h, self.h))
            self.h = None


class PermutationAssembly:
    """The copy-constraint bookkeeping of key generation (`keygen_vk` / `keygen_pk`,
    examples/simple-example.rs:593-594): `copy` joins the cycles of two cells of the permutation columns,
    `sigmas` evaluates the sigma columns on the device (h2a_assembly_*)."""

    def __init__(self, n_cols, k):
        self.lib = load_library()
        self.n_cols, self.k = int(n_cols), int(k)
        self.h = ctypes.c_void_p()
        rc = self.lib.h2a_assembly_new(ctypes.c_uint32(self.n_cols), ctypes.c_uint32(self.k), ctypes.byref(self.h))
        if rc != 0:
            raise H2AError(rc, "h2a_assembly_new")

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.h2a_assembly_free(self.h)
            self.h = None

    def copy(self, col_a, row_a, col_b, row_b):
        rc = self.lib.h2a_assembly_copy(self.h, ctypes.c_uint32(col_a), ctypes.c_uint32(row_a), ctypes.c_uint32(col_b),
                                        ctypes.c_uint32(row_b))
        if rc != 0:
            raise H2AError(rc, "h2a_assembly_copy: cell outside the permutation columns")

    def mapping(self):
        out = np.zeros(self.n_cols << self.k, np.uint32)
        rc = self.lib.h2a_assembly_mapping(self.h, out.ctypes.data_as(ctypes.c_void_p))
        if rc != 0:
            raise H2AError(rc, "h2a_assembly_mapping")
        return out

    def sigmas(self, ctx, omega, delta):
        """n_cols columns of 2^k elements (32-byte Montgomery), the `sigmas` argument of Circuit.set_keys."""
        out = np.zeros((self.n_cols << self.k) * 32, np.uint8)
        ctx._check(self.lib.h2a_assembly_sigmas(ctx.h, self.h, _ptr(_bytes(omega)), _ptr(_bytes(delta)), _ptr(out)))
        return out


class Transcript:
    """Blake2bWrite/Blake2bRead with Challenge255 (src/transcript.rs:58,72,105-107,122-124)."""

    def __init__(self):
        self.lib = load_library()
        self.h = ctypes.c_void_p(self.lib.h2a_transcript_new())

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.h2a_transcript_free(self.h)
            self.h = None

    def common_point(self, p):
        rc = self.lib.h2a_transcript_common_point(self.h, _ptr(_bytes(p)))
        if rc != 0:
            raise H2AError(rc, "common_point: identity or null")

    def common_scalar(self, s):
        rc = self.lib.h2a_transcript_common_scalar(self.h, _ptr(_bytes(s)))
        if rc != 0:
            raise H2AError(rc, "common_scalar")

    def squeeze_challenge(self):
        out = np.zeros(32, np.uint8)
        rc = self.lib.h2a_transcript_squeeze_challenge(self.h, _ptr(out))
        if rc != 0:
            raise H2AError(rc, "squeeze_challenge")
        return out


# ---- free functions named after the dependency functions they replace
def best_multiexp(ctx, coeffs, bases):
    """halo2 `arithmetic::best_multiexp(coeffs, bases)`; `bases` is a Bases handle or a host array."""
    if isinstance(bases, Bases):
        return ctx.msm(bases, coeffs)
    return ctx.msm_adhoc(bases, coeffs)


def best_fft(ctx, a, omega, log_n):
    """halo2 `arithmetic::best_fft(a, omega, log_n)` (returns the transformed copy)."""
    return ctx.ntt(a, log_n, omega)


class EvaluationDomain:
    """Mirror of halo2 `poly::EvaluationDomain` for the methods on the hot path.

    j = cs.degree(); quotient_poly_degree = j - 1; extended_k = smallest with 2^extended_k >= n*(j-1).
    The coset generator is a parameter (the dependency's ZETA is not pinned)."""

    def __init__(self, ctx, j, k, coset_shift):
        self.ctx, self.k = ctx, k
        self.quotient_poly_degree = j - 1
        ext = k
        while (1 << ext) < (1 << k) * (j - 1):
            ext += 1
        self.extended_k = ext
        self.omega = fr_root_of_unity(k)
        self.extended_omega = fr_root_of_unity(ext)
        self.coset_shift = _bytes(coset_shift)

    def get_omega(self):
        return self.omega

    def get_quotient_poly_degree(self):
        return self.quotient_poly_degree

    def lagrange_to_coeff(self, a):
        return self.ctx.ntt(a, self.k, self.omega, inverse=True)

    def coeff_to_lagrange(self, a):
        return self.ctx.ntt(a, self.k, self.omega)

    def coeff_to_extended(self, a):
        return self.ctx.coeff_to_extended(a, self.k, self.extended_k, self.coset_shift)

    def extended_to_coeff(self, a):
        return self.ctx.extended_to_coeff(a, self.extended_k, self.coset_shift)
