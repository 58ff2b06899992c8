"""ctypes loader for the CPU oracle.  TEST INFRASTRUCTURE ONLY (see bn254.hpp).

Only tests/, __graft_entry__.smoke()/build() and bench.py's cpu_baseline /
--impl reference legs may import this module.
"""
import ctypes
import hashlib
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _cpu_tag():
    flags = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    flags = line
                    break
    except OSError:
        pass
    return hashlib.sha1(flags.encode()).hexdigest()[:10]


def lib_path():
    return os.path.join(_HERE, "liboracle-%s.so" % _cpu_tag())


def build(force=False):
    out = lib_path()
    srcs = [os.path.join(_HERE, f) for f in ("oracle.cpp", "bn254.hpp", "Makefile")]
    if force or not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "OUT=" + os.path.basename(out)])
    return out


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.orc_transcript_new.restype = ctypes.c_void_p
        _lib.orc_hw_threads.restype = ctypes.c_int
    return _lib


def hw_threads():
    return load().orc_hw_threads()


def _u8(n):
    return np.zeros(n, dtype=np.uint8)


def _c(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


FQ, FR = 0, 1
OPS = {"add": 0, "sub": 1, "mul": 2, "sqr": 3, "inv": 4, "neg": 5}


def field_op(field, op, a, b=None):
    a = _c(a)
    n = a.size // 32
    out = _u8(32 * n)
    bb = _c(b) if b is not None else None
    load().orc_field_op(field, OPS[op], _p(a), _p(bb), _p(out), ctypes.c_size_t(n))
    return out


def to_mont(field, canonical):
    a = _c(canonical)
    out = _u8(a.size)
    load().orc_to_mont(field, _p(a), _p(out), ctypes.c_size_t(a.size // 32))
    return out


def from_mont(field, mont):
    a = _c(mont)
    out = _u8(a.size)
    load().orc_from_mont(field, _p(a), _p(out), ctypes.c_size_t(a.size // 32))
    return out


def fr_root_of_unity(k):
    out = _u8(32)
    load().orc_fr_root_of_unity(k, _p(out))
    return out


def gen_scalars(seed, n, first=0, threads=None):
    out = _u8(32 * n)
    load().orc_gen_scalars(ctypes.c_uint64(seed), ctypes.c_size_t(first), ctypes.c_size_t(n), _p(out), threads or hw_threads())
    return out


def gen_bases(seed, n, first=0, threads=None):
    out = _u8(64 * n)
    load().orc_gen_bases(ctypes.c_uint64(seed), ctypes.c_size_t(first), ctypes.c_size_t(n), _p(out), threads or hw_threads())
    return out


def g1_on_curve(pts):
    pts = _c(pts)
    return bool(load().orc_g1_on_curve(_p(pts), ctypes.c_size_t(pts.size // 64)))


def g1_add(a, b):
    a, b, out = _c(a), _c(b), _u8(64)
    load().orc_g1_add(_p(a), _p(b), _p(out))
    return out


def g1_mul(a, s):
    a, s, out = _c(a), _c(s), _u8(64)
    load().orc_g1_mul(_p(a), _p(s), _p(out))
    return out


def msm(bases, scalars, threads=None, naive=False):
    bases, scalars = _c(bases), _c(scalars)
    n = scalars.size // 32
    assert bases.size == 64 * n
    out = _u8(64)
    fn = load().orc_msm_naive if naive else load().orc_msm
    fn(_p(bases), _p(scalars), ctypes.c_size_t(n), threads or hw_threads(), _p(out))
    return out


def fft(a, log_n, omega, threads=None):
    a = _c(a).copy()
    load().orc_fft(_p(a), log_n, _p(_c(omega)), threads or hw_threads())
    return a


def ifft(a, log_n, omega_inv, threads=None):
    a = _c(a).copy()
    load().orc_ifft(_p(a), log_n, _p(_c(omega_inv)), threads or hw_threads())
    return a


def coeff_to_extended(coeffs, k, ext_k, shift, threads=None):
    coeffs = _c(coeffs)
    out = _u8(32 << ext_k)
    load().orc_coeff_to_extended(_p(coeffs), k, ext_k, _p(_c(shift)), _p(out), threads or hw_threads())
    return out


def extended_to_coeff(ext, ext_k, shift, threads=None):
    ext = _c(ext).copy()
    load().orc_extended_to_coeff(_p(ext), ext_k, _p(_c(shift)), threads or hw_threads())
    return ext


def blake2b(msg, personal=None, outlen=64):
    msg = np.frombuffer(bytes(msg), dtype=np.uint8) if len(msg) else _u8(0)
    out = _u8(outlen)
    pers = None
    if personal is not None:
        pers = np.frombuffer(bytes(personal).ljust(16, b"\0"), dtype=np.uint8)
    load().orc_blake2b(_p(msg), ctypes.c_size_t(msg.size), _p(pers), ctypes.c_size_t(outlen), _p(out))
    return bytes(out)


def fr_from_bytes_wide(b64):
    a = np.frombuffer(bytes(b64), dtype=np.uint8)
    out = _u8(32)
    load().orc_fr_from_bytes_wide(_p(a), _p(out))
    return out


class Transcript:
    def __init__(self):
        self._t = ctypes.c_void_p(load().orc_transcript_new())

    def __del__(self):
        if self._t:
            load().orc_transcript_free(self._t)
            self._t = None

    def common_point(self, p):
        load().orc_transcript_common_point(self._t, _p(_c(p)))

    def common_scalar(self, s):
        load().orc_transcript_common_scalar(self._t, _p(_c(s)))

    def squeeze(self):
        out = _u8(32)
        load().orc_transcript_squeeze(self._t, _p(out))
        return out


def gwc_accumulate(commitments, rotations, evals, ws, x, u, v, omega, g1):
    commitments, evals, ws = _c(commitments), _c(evals), _c(ws)
    rot = np.ascontiguousarray(rotations, dtype=np.int32)
    out = _u8(256)
    rc = load().orc_gwc_accumulate(_p(commitments), rot.ctypes.data_as(ctypes.c_void_p), _p(evals),
                                   ctypes.c_size_t(rot.size), _p(ws), ctypes.c_size_t(ws.size // 64),
                                   _p(_c(x)), _p(_c(u)), _p(_c(v)), _p(_c(omega)), _p(_c(g1)), _p(out))
    if rc != 0:
        raise ValueError("orc_gwc_accumulate rc=%d" % rc)
    return out


def mulvar_witness_len():
    fn = load().orc_mulvar_witness_len
    fn.restype = ctypes.c_size_t
    return int(fn())


def mulvar_witness(points, scalars, aux, want_cells=True, threads=None):
    """C++ restatement of oracle/mulvar.py: (results m*64, cells m*len*32 or None, status u32[m])"""
    points, scalars, aux = _c(points), _c(scalars), _c(aux)
    m = scalars.size // 32
    assert points.size == 64 * m and aux.size == 64
    res, status = _u8(64 * m), np.zeros(m, np.uint32)
    cells = _u8(32 * mulvar_witness_len() * m) if want_cells else None
    load().orc_mulvar_witness(_p(points), _p(scalars), ctypes.c_size_t(m), _p(aux), _p(res), _p(cells) if want_cells else None,
                              status.ctypes.data_as(ctypes.c_void_p), threads or hw_threads())
    return res, cells, status


def fold_h(h_pieces, xn):
    h = _c(h_pieces)
    out = _u8(64)
    load().orc_fold_h(_p(h), ctypes.c_size_t(h.size // 64), _p(_c(xn)), _p(out))
    return out
