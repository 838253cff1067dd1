"""TEST INFRASTRUCTURE ONLY -- ctypes access to oracle/_build/liboracle.so (the C restatement
of the reference's CPU algorithms, oracle/c/bmpc_oracle.c).  Used by tests/, smoke() and
bench.py's cpu_baseline / --impl reference leg; never by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def _cpu_signature():
    """The library is compiled with -march=native (the timed CPU baseline should be as fast as
    the host allows), so a copy built on another machine is rebuilt before use."""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    import hashlib
                    return hashlib.sha1(line.encode()).hexdigest()
    except OSError:
        pass
    return "unknown"


def build():
    subprocess.run(["make", "-B", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    with open(os.path.join(_HERE, "_build", "cpu.sig"), "w") as f:
        f.write(_cpu_signature())


def load():
    global _lib
    if _lib is not None:
        return _lib
    sig_path = os.path.join(_HERE, "_build", "cpu.sig")
    sig = open(sig_path).read() if os.path.exists(sig_path) else ""
    if not os.path.exists(_PATH) or sig != _cpu_signature():
        build()
    lib = C.CDLL(_PATH)
    vp, sz, i32, u32 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint32
    lib.orc_window_size.restype = u32
    lib.orc_window_size.argtypes = [sz]
    lib.orc_bases_from_uncompressed.restype = vp
    lib.orc_bases_from_uncompressed.argtypes = [i32, vp, sz]
    lib.orc_bases_from_mont.restype = vp
    lib.orc_bases_from_mont.argtypes = [i32, vp, sz]
    lib.orc_bases_free.argtypes = [vp]
    lib.orc_bases_len.restype = sz
    lib.orc_bases_len.argtypes = [vp]
    lib.orc_multiexp.restype = i32
    lib.orc_multiexp.argtypes = [vp, sz, vp, sz, vp, sz, i32, vp]
    lib.orc_multiexp_window.restype = i32
    lib.orc_multiexp_window.argtypes = [vp, sz, vp, sz, sz, i32, vp]
    lib.orc_bases_g1_progression.restype = vp
    lib.orc_bases_g1_progression.argtypes = [sz, vp, vp]
    lib.orc_bases_g1_sequence.restype = vp
    lib.orc_bases_g1_sequence.argtypes = [sz, C.c_uint64]
    lib.orc_naive_multiexp.restype = i32
    lib.orc_naive_multiexp.argtypes = [vp, vp, sz, vp]
    lib.orc_g1_generator_mul.argtypes = [vp, vp]
    lib.orc_fr_dot.argtypes = [vp, vp, sz, vp]
    lib.orc_ntt.restype = i32
    lib.orc_ntt.argtypes = [vp, u32, i32, i32]
    lib.orc_h_coefficients.restype = i32
    lib.orc_h_coefficients.argtypes = [vp, vp, vp, sz, vp, C.POINTER(sz), i32]
    lib.orc_create_proof.restype = i32
    lib.orc_create_proof.argtypes = [vp, vp, vp, vp, sz, vp, sz, vp, sz, vp, vp, vp, vp, vp, i32, vp]
    lib.orc_hardware_threads.restype = i32
    _lib = lib
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def hardware_threads():
    return int(load().orc_hardware_threads())


class CBases:
    def __init__(self, group, handle):
        self.group, self.handle = group, handle

    @staticmethod
    def from_uncompressed(group, data):
        pb = 96 if group == 1 else 192
        buf = np.frombuffer(bytes(data), dtype=np.uint8)
        return CBases(group, load().orc_bases_from_uncompressed(group, _p(buf), buf.size // pb))

    def __len__(self):
        return int(load().orc_bases_len(self.handle))

    def free(self):
        if self.handle:
            load().orc_bases_free(self.handle)
            self.handle = None


def multiexp(bases, start, exps, density_words=None, threads=1):
    """(status, uncompressed bytes) -- status codes as include/bellman_b200.h"""
    exps = np.ascontiguousarray(exps, dtype=np.uint64).reshape(-1, 4)
    n = exps.shape[0]
    out = np.zeros(96 if bases.group == 1 else 192, dtype=np.uint8)
    dw = None if density_words is None else np.ascontiguousarray(density_words, dtype=np.uint64)
    st = load().orc_multiexp(bases.handle, start, _p(exps), n, _p(dw), n if dw is not None else 0, threads, _p(out))
    return st, out.tobytes()


def multiexp_window(bases, start, exps, n_window, threads=1):
    """FullDensity multiexp over `exps` with the window the reference picks for an n_window-entry
    exponent vector (multiexp.rs:267-271): a bounded SAMPLE of a larger workload at that workload's
    per-point cost."""
    exps = np.ascontiguousarray(exps, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros(96 if bases.group == 1 else 192, dtype=np.uint8)
    st = load().orc_multiexp_window(bases.handle, start, _p(exps), exps.shape[0], n_window, threads, _p(out))
    return st, out.tobytes()


def bases_g1_progression(n, first, step):
    """CBases of P_i = (first + i * step) * G (first, step: Python ints < q)"""
    lim = lambda v: np.array([(v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)], dtype=np.uint64)
    return CBases(1, load().orc_bases_g1_progression(n, _p(lim(first)), _p(lim(step))))


def naive_multiexp(bases, exps):
    exps = np.ascontiguousarray(exps, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros(96 if bases.group == 1 else 192, dtype=np.uint8)
    load().orc_naive_multiexp(bases.handle, _p(exps), exps.shape[0], _p(out))
    return out.tobytes()


def g1_generator_mul(k_limbs):
    k = np.ascontiguousarray(k_limbs, dtype=np.uint64).reshape(4)
    out = np.zeros(96, dtype=np.uint8)
    load().orc_g1_generator_mul(_p(k), _p(out))
    return out.tobytes()


def fr_dot(k, s):
    k = np.ascontiguousarray(k, dtype=np.uint64).reshape(-1, 4)
    s = np.ascontiguousarray(s, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros(4, dtype=np.uint64)
    load().orc_fr_dot(_p(k), _p(s), k.shape[0], _p(out))
    return out


def ntt(coeffs_mont, op, threads=1):
    a = np.array(coeffs_mont, dtype=np.uint64, copy=True).reshape(-1, 4)
    logm = int(a.shape[0]).bit_length() - 1
    assert a.shape[0] == 1 << logm
    st = load().orc_ntt(_p(a), logm, op, threads)
    assert st == 0, st
    return a


def h_coefficients(a, b, c, threads=1):
    a, b, c = (np.ascontiguousarray(x, dtype=np.uint64).reshape(-1, 4) for x in (a, b, c))
    n = a.shape[0]
    m = 1
    while m < n:
        m *= 2
    out = np.zeros((max(m, 1), 4), dtype=np.uint64)
    out_len = C.c_size_t()
    st = load().orc_h_coefficients(_p(a), _p(b), _p(c), n, _p(out), C.byref(out_len), threads)
    assert st == 0, st
    return out[: out_len.value]


class CParams(C.Structure):
    _fields_ = [("h", C.c_void_p), ("l", C.c_void_p), ("a", C.c_void_p), ("b_g1", C.c_void_p),
                ("b_g2", C.c_void_p),
                ("alpha_g1", C.c_uint8 * 96), ("beta_g1", C.c_uint8 * 96), ("beta_g2", C.c_uint8 * 192),
                ("delta_g1", C.c_uint8 * 96), ("delta_g2", C.c_uint8 * 192)]


def create_proof(cparams, a, b, c, inputs, aux, a_aux_words, b_in_words, b_aux_words, r_mont, s_mont, threads=1):
    """(status, 192-byte proof)"""
    as4 = lambda x: np.ascontiguousarray(x, dtype=np.uint64).reshape(-1, 4)
    a, b, c, inputs, aux = as4(a), as4(b), as4(c), as4(inputs), as4(aux)
    w = lambda x: np.ascontiguousarray(x, dtype=np.uint64)
    a_aux_words, b_in_words, b_aux_words = w(a_aux_words), w(b_in_words), w(b_aux_words)
    r = np.ascontiguousarray(r_mont, dtype=np.uint64).reshape(4)
    s = np.ascontiguousarray(s_mont, dtype=np.uint64).reshape(4)
    out = np.zeros(192, dtype=np.uint8)
    st = load().orc_create_proof(C.byref(cparams), _p(a), _p(b), _p(c), a.shape[0], _p(inputs), inputs.shape[0],
                                 _p(aux), aux.shape[0], _p(a_aux_words), _p(b_in_words), _p(b_aux_words),
                                 _p(r), _p(s), threads, _p(out))
    return st, out.tobytes()
