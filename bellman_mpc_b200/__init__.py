"""bellman_mpc_b200 -- host-side mirror of the reference's hot-path interface over the
C ABI of libbellman_b200.so (include/bellman_b200.h).

Names, argument meaning and error behaviour follow the Rust reference:

  Worker / Waiter .......................... src/multicore.rs:21-118
  FullDensity / DensityTracker ............. src/multiexp.rs:88-157
  multiexp(pool, bases, density, exponents)  src/multiexp.rs:254-281
  EvaluationDomain ......................... src/domain.rs:21-189
  Parameters / create_proof ................ src/groth16/mod.rs:224-247, prover.rs:176-350
  SynthesisError variants .................. src/lib.rs:355-370

Data crosses the boundary as numpy arrays of little-endian u64 limbs, i.e. exactly the bytes
the Rust types hold (FieldBits<[u64;4]> canonical exponents; Scalar<Fr> Montgomery limbs).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import (COSET_FFT, FFT, FORM_MONT_XY, FORM_UNCOMPRESSED_BE, G1, G2, ICOSET_FFT, IFFT)

FR_MODULUS = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
_FR_R = (1 << 256) % FR_MODULUS


# ------------------------------------------------------------------------------ errors
class SynthesisError(Exception):
    """src/lib.rs:355-370"""


class UnexpectedIdentity(SynthesisError):
    pass


class IoError(SynthesisError):
    pass


class UnexpectedEof(IoError):
    pass


class PolynomialDegreeTooLarge(SynthesisError):
    pass


class InvalidData(IoError):
    """io::ErrorKind::InvalidData from Parameters::read ("invalid G1", "point at infinity")"""


def _raise(status, ctx=None, msg_override=None):
    if status == _lib.OK:
        return
    if msg_override is not None and status in (_lib.ERR_CUDA, _lib.ERR_INVALID, _lib.ERR_INVALID_DATA):
        if status == _lib.ERR_CUDA:
            raise IoError(f"CUDA error: {msg_override}")
        if status == _lib.ERR_INVALID_DATA:
            raise InvalidData(msg_override)
        raise ValueError(f"invalid argument ({msg_override})")
    if status == _lib.ERR_UNEXPECTED_IDENTITY:
        raise UnexpectedIdentity()
    if status == _lib.ERR_UNEXPECTED_EOF:
        raise UnexpectedEof("expected more bases from source")
    if status == _lib.ERR_DEGREE_TOO_LARGE:
        raise PolynomialDegreeTooLarge()
    if status == _lib.ERR_INVALID_DATA:
        raise InvalidData(_lib.load().bmpc_last_error(ctx).decode() if ctx is not None else "invalid data")
    if status == _lib.ERR_LENGTH_MISMATCH:
        raise AssertionError("length mismatch")        # the reference panics (assert!)
    msg = ""
    if ctx is not None:
        msg = _lib.load().bmpc_last_error(ctx).decode()
    if status == _lib.ERR_CUDA:
        raise IoError(f"CUDA error: {msg}")
    raise ValueError(f"invalid argument ({msg})")


# ---------------------------------------------------------------------- limb helpers
def ints_to_limbs(vals, words=4):
    """list of Python ints -> (n, words) u64 little-endian limbs"""
    out = np.empty((len(vals), words), dtype=np.uint64)
    mask = (1 << 64) - 1
    for i, v in enumerate(vals):
        for j in range(words):
            out[i, j] = (v >> (64 * j)) & mask
    return out


def limbs_to_ints(arr):
    arr = np.asarray(arr, dtype=np.uint64)
    return [sum(int(x) << (64 * j) for j, x in enumerate(row)) for row in arr.reshape(-1, arr.shape[-1])]


def fr_to_mont(vals):
    """canonical ints -> Montgomery limbs (what Scalar<Fr> holds)"""
    return ints_to_limbs([v * _FR_R % FR_MODULUS for v in vals])


def fr_from_mont(arr):
    rinv = pow(_FR_R, -1, FR_MODULUS)
    return [v * rinv % FR_MODULUS for v in limbs_to_ints(arr)]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# ------------------------------------------------------------------- Worker / Waiter
class Worker:
    """multicore.rs:21-91.  Here the pool is one GPU context."""

    def __init__(self, device=0):
        self._lib = _lib.load()
        h = C.c_void_p()
        st = self._lib.bmpc_ctx_create(int(device), C.byref(h))
        if st != _lib.OK:
            raise IoError(f"bmpc_ctx_create(device={device}) failed with status {st}: "
                          "no usable CUDA device (there is no CPU fallback)")
        self.ctx = h
        self.device = device

    def set_tuning(self, msm_window_bits=0, ntt_max_deg=0):
        _raise(self._lib.bmpc_ctx_set_tuning(self.ctx, msm_window_bits, ntt_max_deg), self.ctx)

    def reload_env(self):
        """re-read the BMPC_* tuning variables (they are parsed once per context otherwise)"""
        _raise(self._lib.bmpc_ctx_reload_env(self.ctx), self.ctx)

    def launch_count(self):
        return int(self._lib.bmpc_ctx_launch_count(self.ctx))

    def close(self):
        if self.ctx:
            self._lib.bmpc_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Waiter:
    """multicore.rs:93-118: `wait()` returns the Result (raises the SynthesisError).  A waiter made by
    `multiexp` holds a bmpc_waiter: the multiexp is in flight on a lane of the context and wait()
    blocks for it (bmpc_waiter_wait)."""

    def __init__(self, value=None, error=None, pending=None):
        self._value, self._error, self._pending = value, error, pending

    @staticmethod
    def done(value):
        return Waiter(value=value)

    def wait(self):
        if self._pending is not None:
            pool, handle, out, keep = self._pending
            self._pending = None
            st = pool._lib.bmpc_waiter_wait(handle, _ptr(out))
            del keep                                   # scalars / density words had to outlive the call
            try:
                _raise(st, pool.ctx)
                self._value = out.tobytes()
            except SynthesisError as e:
                self._error = e
        if self._error is not None:
            raise self._error
        return self._value

    def __del__(self):
        # a waiter that is dropped still has to give its lane back
        try:
            if self._pending is not None:
                self.wait()
        except Exception:
            pass


class MultiWorker:
    """All GPUs of a node behind one Worker (SURVEY 8b `bmpc_ctx_create(devices, n)`): `multiexp` and
    `create_proof` stay single calls; the library shards them over the devices (csrc/multi.cu)."""

    def __init__(self, devices):
        self._lib = _lib.load()
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        st = self._lib.bmpc_multi_create(devs, len(devices), C.byref(h))
        if st != _lib.OK:
            raise IoError(f"bmpc_multi_create({list(devices)}) failed with status {st}: "
                          "no usable CUDA device (there is no CPU fallback)")
        self.handle = h
        self.devices = list(devices)

    def __len__(self):
        return int(self._lib.bmpc_multi_size(self.handle))

    def last_error(self):
        return self._lib.bmpc_multi_last_error(self.handle).decode()

    def rank_ctx(self, rank):
        return C.c_void_p(self._lib.bmpc_multi_ctx(self.handle, rank))

    def close(self):
        if self.handle:
            self._lib.bmpc_multi_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiBases:
    """A base vector split evenly over the devices of a MultiWorker (bmpc_multi_bases)."""

    def __init__(self, worker, group, handle):
        self.worker, self.group, self.handle = worker, group, handle
        self._lib = worker._lib

    @staticmethod
    def from_uncompressed(worker, group, data, n=None):
        pb = 96 if group == G1 else 192
        buf = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        if n is None:
            n = buf.size // pb
        h = C.c_void_p()
        _raise(worker._lib.bmpc_multi_bases_register(worker.handle, group, _ptr(buf), n, pb, FORM_UNCOMPRESSED_BE,
                                                     C.byref(h)), None, worker.last_error())
        return MultiBases(worker, group, h)

    def precompute(self, window_bits=0):
        _raise(self._lib.bmpc_multi_bases_precompute(self.worker.handle, self.handle, window_bits), None,
               self.worker.last_error())
        return self

    def __len__(self):
        return int(self._lib.bmpc_multi_bases_len(self.handle))

    def free(self):
        if self.handle:
            self._lib.bmpc_multi_bases_free(self.worker.handle, self.handle)
            self.handle = None


class MultiParameters:
    """groth16/mod.rs:224-247 with every query vector split over the devices of a MultiWorker."""

    def __init__(self, worker, h, l, a, b_g1, b_g2, alpha_g1, beta_g1, beta_g2, delta_g1, delta_g2):
        self.worker = worker
        self.h, self.l, self.a, self.b_g1, self.b_g2 = h, l, a, b_g1, b_g2
        self.alpha_g1, self.beta_g1, self.beta_g2 = bytes(alpha_g1), bytes(beta_g1), bytes(beta_g2)
        self.delta_g1, self.delta_g2 = bytes(delta_g1), bytes(delta_g2)

    def _struct(self):
        p = _lib.MultiParams()
        p.h, p.l, p.a, p.b_g1, p.b_g2 = (self.h.handle, self.l.handle, self.a.handle, self.b_g1.handle,
                                         self.b_g2.handle)
        for name in ("alpha_g1", "beta_g1", "beta_g2", "delta_g1", "delta_g2"):
            C.memmove(getattr(p, name), getattr(self, name), len(getattr(self, name)))
        return p

    def free(self):
        for b in (self.h, self.l, self.a, self.b_g1, self.b_g2):
            b.free()


# ----------------------------------------------------------------------------- bases
class Bases:
    """Arc<Vec<G::Affine>> resident in HBM; `(bases, offset)` is the reference's SourceBuilder
    (multiexp.rs:45-51)."""

    def __init__(self, worker, handle):
        self.worker, self.handle = worker, handle
        self._lib = worker._lib

    @staticmethod
    def from_uncompressed(worker, group, data, n=None):
        pb = 96 if group == G1 else 192
        buf = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        if n is None:
            n = buf.size // pb
        h = C.c_void_p()
        _raise(worker._lib.bmpc_bases_register(worker.ctx, group, _ptr(buf), n, pb, FORM_UNCOMPRESSED_BE,
                                               C.byref(h)), worker.ctx)
        return Bases(worker, h)

    @staticmethod
    def from_mont(worker, group, limbs):
        pb = 96 if group == G1 else 192
        arr = np.ascontiguousarray(limbs, dtype=np.uint64)
        n = arr.size * 8 // pb
        h = C.c_void_p()
        _raise(worker._lib.bmpc_bases_register(worker.ctx, group, _ptr(arr), n, pb, FORM_MONT_XY,
                                               C.byref(h)), worker.ctx)
        return Bases(worker, h)

    @staticmethod
    def fixed_base_mul(worker, group, base_uncompressed, scalars):
        """out[i] = base * scalars[i]  (canonical (n,4) u64)"""
        sc = np.ascontiguousarray(scalars, dtype=np.uint64)
        base = np.frombuffer(bytes(base_uncompressed), dtype=np.uint8)
        h = C.c_void_p()
        _raise(worker._lib.bmpc_fixed_base_mul(worker.ctx, group, _ptr(base), _ptr(sc), sc.shape[0], 0,
                                               C.byref(h)), worker.ctx)
        return Bases(worker, h)

    def scalar_mul(self, scalars, per_element=True):
        """mpc.rs:647-706: every element times its own scalar, or all times one scalar."""
        sc = np.ascontiguousarray(scalars, dtype=np.uint64)
        h = C.c_void_p()
        _raise(self._lib.bmpc_batch_scalar_mul(self.worker.ctx, self.handle, _ptr(sc), 1 if per_element else 0,
                                               C.byref(h)), self.worker.ctx)
        return Bases(self.worker, h)

    def precompute(self, window_bits=0):
        """window tables 2^(c w) * P_i in HBM (one bucket set, no doubling fold); returns self"""
        _raise(self._lib.bmpc_bases_precompute(self.worker.ctx, self.handle, window_bits), self.worker.ctx)
        return self

    @property
    def group(self):
        return self._lib.bmpc_bases_group(self.handle)

    def __len__(self):
        return int(self._lib.bmpc_bases_len(self.handle))

    def read(self, start=0, count=None):
        if count is None:
            count = len(self) - start
        pb = 96 if self.group == G1 else 192
        out = np.empty(count * pb, dtype=np.uint8)
        _raise(self._lib.bmpc_bases_read(self.worker.ctx, self.handle, start, count, _ptr(out)), self.worker.ctx)
        return out.tobytes()

    def free(self):
        if self.handle:
            self._lib.bmpc_bases_free(self.worker.ctx, self.handle)
            self.handle = None


def list_mul_matrix(list_g1, list_g2, matrix):
    """mpc.rs:416-457 `list_mul_matrix(list_g1, list_g2, matrix) -> (Vec<G1>, Vec<G2>)`:
    result[i] = sum_j list[matrix[i][j].1] * matrix[i][j].0 for the rows before the first empty
    one, identity elsewhere; both results have the lists' length.  `matrix` is the reference's
    `Vec<Vec<(Fr, usize)>>` with canonical integers for Fr.  The reference panics on an index out
    of range (matrix taller than the list, column >= list length): AssertionError here."""
    row_ptr = np.zeros(len(matrix) + 1, dtype=np.uint64)
    for i, row in enumerate(matrix):
        row_ptr[i + 1] = row_ptr[i] + len(row)
    cols = np.array([idx for row in matrix for (_, idx) in row], dtype=np.int64)
    if cols.size and (cols.min() < 0 or cols.max() >= (1 << 32)):
        raise AssertionError("index out of bounds")
    cols = np.ascontiguousarray(cols, dtype=np.uint32)
    coeffs = ints_to_limbs([cf for row in matrix for (cf, _) in row]) if cols.size else np.zeros((0, 4), np.uint64)
    out = []
    for lst in (list_g1, list_g2):
        h = C.c_void_p()
        _raise(lst._lib.bmpc_list_mul_matrix(lst.worker.ctx, lst.handle, _ptr(row_ptr), _ptr(cols) if cols.size else None,
                                             _ptr(coeffs) if cols.size else None, len(matrix), C.byref(h)),
               lst.worker.ctx)
        out.append(Bases(lst.worker, h))
    return tuple(out)


# --------------------------------------------------------------------------- density
def fold_vectors(pool, vectors, rho):
    """The GPU half of a ceremony check batched by random linear combination (the reference checks
    every element of a contribution with two pairings, groth16/mpc.rs:806-862,1065-1131): for each
    resident vector V returns sum_i rho_i V_i (uncompressed bytes), all multiexps in flight together.
    The caller draws `rho` (n scalars, canonical (n, 4) u64; 128 random bits each are enough) AFTER it
    has received the vectors, and finishes with two pairings per pair of folded points."""
    exps = np.ascontiguousarray(rho, dtype=np.uint64).reshape(-1, 4)
    waiters = [multiexp(pool, (v, 0), FullDensity(), exps[:len(v)]) for v in vectors]
    return [w.wait() for w in waiters]


class FullDensity:
    """multiexp.rs:95-114"""

    def get_query_size(self):
        return None

    def words(self):
        return None


class DensityTracker:
    """multiexp.rs:116-157.  Bits are kept in a numpy array; `words()` is the raw
    BitVec<Lsb0, usize> storage handed to the C ABI (cached until the next mutation)."""

    def __init__(self):
        self._bits = np.zeros(0, dtype=bool)
        self._n = 0
        self._words = None

    @property
    def bv(self):
        return self._bits[: self._n]

    def add_element(self):
        if self._n == self._bits.size:
            grown = np.zeros(max(64, 2 * self._bits.size), dtype=bool)
            grown[: self._n] = self._bits[: self._n]
            self._bits = grown
        self._bits[self._n] = False
        self._n += 1
        self._words = None

    def inc(self, idx):
        if idx >= self._n:
            raise IndexError(idx)
        self._bits[idx] = True
        self._words = None

    def get_total_density(self):
        return int(self.bv.sum())

    def get_query_size(self):
        return self._n

    def words(self):
        if self._words is None:
            bits = self.bv.astype(np.uint8)
            pad = (-len(bits)) % 64
            if pad or len(bits) == 0:
                bits = np.concatenate([bits, np.zeros(pad if len(bits) else 64, dtype=np.uint8)])
            self._words = np.packbits(bits, bitorder="little").view(np.uint64).copy()
        return self._words

    @staticmethod
    def from_bits(bits):
        d = DensityTracker()
        d._bits = np.array(bits, dtype=bool).reshape(-1)
        d._n = d._bits.size
        return d


# -------------------------------------------------------------------------- multiexp
def multiexp(pool, bases, density_map, exponents):
    """multiexp.rs:254-281.  bases = (Bases, start_index); exponents = (n, 4) u64 canonical.
    Returns a Waiter whose wait() yields the uncompressed affine sum (bytes) or raises."""
    src, start = bases
    exps = np.ascontiguousarray(exponents, dtype=np.uint64).reshape(-1, 4)
    n = exps.shape[0]
    qs = density_map.get_query_size()
    if qs is not None:
        assert qs == n                                             # multiexp.rs:273-278
    words = density_map.words()
    pb = 96 if src.group == G1 else 192
    out = np.zeros(pb, dtype=np.uint8)
    if isinstance(pool, MultiWorker):
        st = pool._lib.bmpc_multi_multiexp(pool.handle, src.handle, start, _ptr(exps), n, _ptr(words),
                                           n if words is not None else 0, _ptr(out))
        try:
            _raise(st, None, pool.last_error())
        except SynthesisError as e:
            return Waiter(error=e)
        return Waiter.done(out.tobytes())
    h = C.c_void_p()
    st = pool._lib.bmpc_multiexp_async(pool.ctx, src.handle, start, _ptr(exps), n, _ptr(words),
                                       n if words is not None else 0, C.byref(h))
    try:
        _raise(st, pool.ctx)
    except SynthesisError as e:
        return Waiter(error=e)
    return Waiter(pending=(pool, h, out, (exps, words)))


# ------------------------------------------------------------------ EvaluationDomain
class EvaluationDomain:
    """domain.rs:21-189 for Scalar<Fr>; coefficients live in HBM."""

    def __init__(self, worker, handle):
        self.worker, self.handle = worker, handle
        self._lib = worker._lib

    @staticmethod
    def from_coeffs(worker, coeffs):
        """coeffs: (len, 4) u64 Montgomery limbs"""
        arr = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
        h = C.c_void_p()
        _raise(worker._lib.bmpc_domain_from_coeffs(worker.ctx, _ptr(arr), arr.shape[0], C.byref(h)), worker.ctx)
        return EvaluationDomain(worker, h)

    def __len__(self):
        return int(self._lib.bmpc_domain_len(self.handle))

    @property
    def exp(self):
        return int(self._lib.bmpc_domain_exp(self.handle))

    def into_coeffs(self):
        out = np.empty((len(self), 4), dtype=np.uint64)
        _raise(self._lib.bmpc_domain_into_coeffs(self.worker.ctx, self.handle, _ptr(out)), self.worker.ctx)
        return out

    def _t(self, op):
        _raise(self._lib.bmpc_domain_transform(self.worker.ctx, self.handle, op, None), self.worker.ctx)

    def fft(self, worker=None):
        self._t(FFT)

    def ifft(self, worker=None):
        self._t(IFFT)

    def coset_fft(self, worker=None):
        self._t(COSET_FFT)

    def icoset_fft(self, worker=None):
        self._t(ICOSET_FFT)

    def distribute_powers(self, worker, g_mont):
        g = np.ascontiguousarray(g_mont, dtype=np.uint64).reshape(4)
        _raise(self._lib.bmpc_domain_distribute_powers(self.worker.ctx, self.handle, _ptr(g), None), self.worker.ctx)

    def z(self, tau_mont):
        t = np.ascontiguousarray(tau_mont, dtype=np.uint64).reshape(4)
        out = np.empty(4, dtype=np.uint64)
        _raise(self._lib.bmpc_domain_z(self.worker.ctx, self.handle, _ptr(t), _ptr(out)), self.worker.ctx)
        return out

    def divide_by_z_on_coset(self, worker=None):
        _raise(self._lib.bmpc_domain_divide_by_z_on_coset(self.worker.ctx, self.handle, None), self.worker.ctx)

    def mul_assign(self, worker, other):
        _raise(self._lib.bmpc_domain_mul_assign(self.worker.ctx, self.handle, other.handle, None), self.worker.ctx)

    def sub_assign(self, worker, other):
        _raise(self._lib.bmpc_domain_sub_assign(self.worker.ctx, self.handle, other.handle, None), self.worker.ctx)

    def free(self):
        if self.handle:
            self._lib.bmpc_domain_free(self.worker.ctx, self.handle)
            self.handle = None


def h_coefficients(worker, a, b, c):
    """prover.rs:210-231: (len,4) Montgomery evaluations -> (m-1, 4) canonical quotient scalars"""
    a, b, c = (np.ascontiguousarray(x, dtype=np.uint64).reshape(-1, 4) for x in (a, b, c))
    n = a.shape[0]
    m = 1
    while m < n:
        m *= 2
    out = np.empty((max(m - 1, 0), 4), dtype=np.uint64)
    out_len = C.c_size_t()
    _raise(worker._lib.bmpc_h_coefficients(worker.ctx, _ptr(a), _ptr(b), _ptr(c), n,
                                           _ptr(out) if out.size else _ptr(np.empty((1, 4), dtype=np.uint64)),
                                           C.byref(out_len)), worker.ctx)
    return out[: out_len.value]


# --------------------------------------------------------------------------- groth16
class Parameters:
    """groth16/mod.rs:224-247 with the query vectors resident in HBM.  vk_* are uncompressed
    big-endian encodings."""

    def __init__(self, worker, h, l, a, b_g1, b_g2, alpha_g1, beta_g1, beta_g2, delta_g1, delta_g2):
        self.worker = worker
        self.h, self.l, self.a, self.b_g1, self.b_g2 = h, l, a, b_g1, b_g2
        self.alpha_g1, self.beta_g1, self.beta_g2 = bytes(alpha_g1), bytes(beta_g1), bytes(beta_g2)
        self.delta_g1, self.delta_g2 = bytes(delta_g1), bytes(delta_g2)

    @staticmethod
    def read(worker, data, checked):
        """Parameters::read(reader, checked) (groth16/mod.rs:292-400): decode + validate on the GPU,
        query vectors stay resident."""
        buf = np.frombuffer(bytes(data), dtype=np.uint8)
        pf = _lib.ParametersFile()
        _raise(worker._lib.bmpc_params_read(worker.ctx, _ptr(buf), buf.size, 1 if checked else 0, C.byref(pf)),
               worker.ctx)
        mk = lambda h: Bases(worker, C.c_void_p(h))
        out = Parameters(worker, mk(pf.p.h), mk(pf.p.l), mk(pf.p.a), mk(pf.p.b_g1), mk(pf.p.b_g2),
                         bytes(pf.p.alpha_g1), bytes(pf.p.beta_g1), bytes(pf.p.beta_g2), bytes(pf.p.delta_g1),
                         bytes(pf.p.delta_g2))
        out.gamma_g2 = bytes(pf.gamma_g2)
        out.ic = mk(pf.ic)
        return out

    def write(self):
        """Parameters::write (groth16/mod.rs:261-290)"""
        pf = _lib.ParametersFile()
        pf.p = self._struct()
        C.memmove(pf.gamma_g2, self.gamma_g2, 192)
        pf.ic = self.ic.handle
        n = C.c_size_t()
        self.worker._lib.bmpc_params_write(self.worker.ctx, C.byref(pf), None, 0, C.byref(n))
        out = np.empty(n.value, dtype=np.uint8)
        _raise(self.worker._lib.bmpc_params_write(self.worker.ctx, C.byref(pf), _ptr(out), n.value, C.byref(n)),
               self.worker.ctx)
        return out.tobytes()

    def free(self):
        for b in (self.h, self.l, self.a, self.b_g1, self.b_g2, getattr(self, "ic", None)):
            if b is not None:
                b.free()

    def _struct(self):
        p = _lib.Params()
        p.h, p.l, p.a, p.b_g1, p.b_g2 = (self.h.handle, self.l.handle, self.a.handle, self.b_g1.handle,
                                         self.b_g2.handle)
        for name in ("alpha_g1", "beta_g1", "beta_g2", "delta_g1", "delta_g2"):
            C.memmove(getattr(p, name), getattr(self, name), len(getattr(self, name)))
        return p


class ProvingAssignment:
    """prover.rs:55-69 after synthesis: evaluations a, b, c and assignments as Montgomery limbs,
    densities as DensityTracker."""

    def __init__(self, a, b, c, input_assignment, aux_assignment, a_aux_density, b_input_density,
                 b_aux_density):
        as4 = lambda x: np.ascontiguousarray(x, dtype=np.uint64).reshape(-1, 4)
        self.a, self.b, self.c = as4(a), as4(b), as4(c)
        self.input_assignment, self.aux_assignment = as4(input_assignment), as4(aux_assignment)
        self.a_aux_density, self.b_input_density, self.b_aux_density = a_aux_density, b_input_density, b_aux_density


def create_proof(assignment, params, r_mont, s_mont):
    """prover.rs:206-350 (everything after synthesis).  Returns the 192-byte proof."""
    w = params.worker
    asg = assignment
    assert asg.a.shape == asg.b.shape == asg.c.shape
    assert asg.a_aux_density.get_query_size() == asg.aux_assignment.shape[0]
    assert asg.b_aux_density.get_query_size() == asg.aux_assignment.shape[0]
    assert asg.b_input_density.get_query_size() == asg.input_assignment.shape[0]
    s = _lib.Assignment()
    keep = [asg.a_aux_density.words(), asg.b_input_density.words(), asg.b_aux_density.words()]
    s.a, s.b, s.c = _ptr(asg.a), _ptr(asg.b), _ptr(asg.c)
    s.num_constraints = asg.a.shape[0]
    s.input_assignment, s.num_inputs = _ptr(asg.input_assignment), asg.input_assignment.shape[0]
    s.aux_assignment, s.num_aux = _ptr(asg.aux_assignment), asg.aux_assignment.shape[0]
    s.a_aux_density, s.b_input_density, s.b_aux_density = (_ptr(k) for k in keep)
    p = params._struct()
    r = np.ascontiguousarray(r_mont, dtype=np.uint64).reshape(4)
    sv = np.ascontiguousarray(s_mont, dtype=np.uint64).reshape(4)
    out = np.zeros(192, dtype=np.uint8)
    if isinstance(w, MultiWorker):
        _raise(w._lib.bmpc_multi_create_proof(w.handle, C.byref(p), C.byref(s), _ptr(r), _ptr(sv), _ptr(out)),
               None, w.last_error())
    else:
        _raise(w._lib.bmpc_create_proof(w.ctx, C.byref(p), C.byref(s), _ptr(r), _ptr(sv), _ptr(out)), w.ctx)
    return out.tobytes()


class CsrMatrix:
    """R1CS matrix in CSR form (bmpc_csr): rows of (column, Montgomery coefficient)."""

    def __init__(self, row_ptr, col, coeff_mont):
        self.row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint32)
        self.col = np.ascontiguousarray(col, dtype=np.uint32)
        self.coeff = np.ascontiguousarray(coeff_mont, dtype=np.uint64).reshape(-1, 4)
        assert self.row_ptr[-1] == len(self.col) == self.coeff.shape[0]

    @staticmethod
    def from_rows(rows):
        """rows: list of lists of (column, canonical int coefficient)"""
        row_ptr, col, coeff = [0], [], []
        for r in rows:
            for c, v in r:
                col.append(c)
                coeff.append(v % FR_MODULUS)
            row_ptr.append(len(col))
        return CsrMatrix(row_ptr, col, fr_to_mont(coeff) if coeff else np.zeros((0, 4), dtype=np.uint64))

    @property
    def num_rows(self):
        return len(self.row_ptr) - 1

    def _struct(self):
        s = _lib.Csr()
        s.row_ptr, s.col, s.coeff = _ptr(self.row_ptr), _ptr(self.col), _ptr(self.coeff)
        s.num_rows, s.nnz = self.num_rows, len(self.col)
        return s


def r1cs_eval(worker, A, B, C_, input_assignment, aux_assignment):
    """ProvingAssignment::enforce over the whole system + create_proof's input rows
    (prover.rs:19-53,100-138,202-204) -> ProvingAssignment ready for create_proof."""
    as4 = lambda x: np.ascontiguousarray(x, dtype=np.uint64).reshape(-1, 4)
    inp, aux = as4(input_assignment), as4(aux_assignment)
    ni, na = inp.shape[0], aux.shape[0]
    total = A.num_rows + ni
    a, b, c = (np.empty((total, 4), dtype=np.uint64) for _ in range(3))
    da, dbi, dba = (np.zeros(max(1, (n + 63) // 64), dtype=np.uint64) for n in (na, ni, na))
    sa, sb, sc = A._struct(), B._struct(), C_._struct()
    _raise(worker._lib.bmpc_r1cs_eval(worker.ctx, C.byref(sa), C.byref(sb), C.byref(sc), ni, na, _ptr(inp), _ptr(aux),
                                      _ptr(a), _ptr(b), _ptr(c), _ptr(da), _ptr(dbi), _ptr(dba)), worker.ctx)
    unpack = lambda w, n: DensityTracker.from_bits(np.unpackbits(w.view(np.uint8), bitorder="little")[:n])
    return ProvingAssignment(a, b, c, inp, aux, unpack(da, na), unpack(dbi, ni), unpack(dba, na))


def generate_parameters(worker, At, Bt, Ct, num_inputs, num_aux, num_constraints, g1, g2, alpha, beta, gamma, delta,
                        tau):
    """generate_parameters, upstream semantics (generator.rs:241-634).  At/Bt/Ct: transposed R1CS
    (one row per variable); scalars canonical ints; g1/g2 uncompressed bytes."""
    pf = _lib.ParametersFile()
    sa, sb, sc = At._struct(), Bt._struct(), Ct._struct()
    g1b, g2b = np.frombuffer(bytes(g1), dtype=np.uint8), np.frombuffer(bytes(g2), dtype=np.uint8)
    ks = [fr_to_mont([k])[0] for k in (alpha, beta, gamma, delta, tau)]
    _raise(worker._lib.bmpc_generate_parameters(worker.ctx, C.byref(sa), C.byref(sb), C.byref(sc), num_inputs, num_aux,
                                                num_constraints, _ptr(g1b), _ptr(g2b), *(_ptr(k) for k in ks),
                                                C.byref(pf)), worker.ctx)
    mk = lambda h: Bases(worker, C.c_void_p(h))
    out = Parameters(worker, mk(pf.p.h), mk(pf.p.l), mk(pf.p.a), mk(pf.p.b_g1), mk(pf.p.b_g2),
                     bytes(pf.p.alpha_g1), bytes(pf.p.beta_g1), bytes(pf.p.beta_g2), bytes(pf.p.delta_g1),
                     bytes(pf.p.delta_g2))
    out.gamma_g2 = bytes(pf.gamma_g2)
    out.ic = mk(pf.ic)
    return out


def create_random_proof(assignment, params, rng=None):
    """prover.rs:158-173: the fork ignores the RNG and uses r = 27134, s = 17146."""
    return create_proof(assignment, params, fr_to_mont([27134])[0], fr_to_mont([17146])[0])
