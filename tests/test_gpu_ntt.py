"""GPU parity: EvaluationDomain transforms vs the oracle restatement of src/domain.rs.
Mirrors the reference's own tests polynomial_arith (domain.rs:376-425), fft_composition
(:427-463) and parallel_fft_consistency (:465-498), plus bit-exactness against serial_fft."""
import random

import numpy as np
import pytest

import bellman_mpc_b200 as bm
from oracle import domain as odomain
from oracle import fields

pytestmark = pytest.mark.gpu
F = fields.Fr
Q = F.p


def _rand(n, seed):
    rng = random.Random(seed)
    return [rng.randrange(Q) for _ in range(n)]


@pytest.mark.parametrize("logn", list(range(0, 13)) + [14])
def test_transforms_bit_exact(worker, logn):
    n = 1 << logn
    coeffs = _rand(n, 100 + logn)
    for op, name in [(bm.FFT, "fft"), (bm.IFFT, "ifft"), (bm.COSET_FFT, "coset_fft"), (bm.ICOSET_FFT, "icoset_fft")]:
        d = bm.EvaluationDomain.from_coeffs(worker, bm.fr_to_mont(coeffs))
        getattr(d, name)(worker)
        got = bm.fr_from_mont(d.into_coeffs())
        d.free()
        o = odomain.EvaluationDomain(F, coeffs)
        getattr(o, name)()
        assert got == o.coeffs, (logn, name)


def _rand_mont(m, seed):
    """m x 4 u64, uniform below 2^254 (< q): valid fully reduced Montgomery residues"""
    rs = np.random.RandomState(seed)
    a = rs.randint(0, 1 << 63, size=(m, 4), dtype=np.int64).astype(np.uint64)
    a[:, 3] >>= np.uint64(1)
    return a


@pytest.mark.parametrize("logn", [16, 18, 20, 22, 24, 26])
def test_transforms_bit_exact_sweep(worker, logn):
    """BASELINE config #3: fft / ifft / coset_fft / icoset_fft over 2^16 .. 2^26, every limb of every
    coefficient equal to the C restatement of domain.rs:81-125,261-372 (oracle/c, itself pinned to the
    Python oracle and through it to the reference's known answers).  2^26 is the 4-pass transform
    with the two-level twiddle lookup (direct power tables end at 2^24)."""
    import ctypes as C
    from oracle import cref
    m = 1 << logn
    coeffs = _rand_mont(m, 500 + logn)
    threads = cref.hardware_threads()
    lib = worker._lib
    for op in (bm.FFT, bm.IFFT, bm.COSET_FFT, bm.ICOSET_FFT):
        want = cref.ntt(coeffs, op, threads=threads)
        got = coeffs.copy()
        rc = lib.bmpc_ntt(worker.ctx, got.ctypes.data_as(C.c_void_p), logn, op)
        assert rc == 0
        assert np.array_equal(got, want), (logn, op, int((got != want).any(axis=1).sum()))
        del want, got


@pytest.mark.parametrize("maxdeg", [3, 5, 10])
def test_pass_split_consistency(worker, maxdeg):
    """parallel_fft_consistency analogue: the result must not depend on how the transform is
    split into passes (the reference: on log_cpus)."""
    n = 1 << 11
    coeffs = _rand(n, 7)
    o = odomain.EvaluationDomain(F, coeffs)
    o.fft()
    worker.set_tuning(0, maxdeg)
    try:
        d = bm.EvaluationDomain.from_coeffs(worker, bm.fr_to_mont(coeffs))
        d.fft(worker)
        assert bm.fr_from_mont(d.into_coeffs()) == o.coeffs
        d.free()
    finally:
        worker.set_tuning(0, 0)


def test_fft_composition(worker):
    """domain.rs:427-463"""
    for logn in range(0, 10):
        n = 1 << logn
        coeffs = _rand(n, 300 + logn)
        for f, g in [("ifft", "fft"), ("fft", "ifft"), ("icoset_fft", "coset_fft"), ("coset_fft", "icoset_fft")]:
            d = bm.EvaluationDomain.from_coeffs(worker, bm.fr_to_mont(coeffs))
            getattr(d, f)(worker)
            getattr(d, g)(worker)
            assert bm.fr_from_mont(d.into_coeffs()) == coeffs, (logn, f, g)
            d.free()


def test_polynomial_arith(worker):
    """domain.rs:376-425: fft * fft -> mul_assign -> ifft equals the schoolbook product"""
    rng = random.Random(5)
    for na, nb in [(1, 1), (3, 5), (17, 40), (69, 69)]:
        a = [rng.randrange(Q) for _ in range(na)]
        b = [rng.randrange(Q) for _ in range(nb)]
        naive = [0] * (na + nb)
        for i, x in enumerate(a):
            for j, y in enumerate(b):
                naive[i + j] = (naive[i + j] + x * y) % Q
        size = na + nb
        da = bm.EvaluationDomain.from_coeffs(worker, bm.fr_to_mont(a + [0] * (size - na)))
        db = bm.EvaluationDomain.from_coeffs(worker, bm.fr_to_mont(b + [0] * (size - nb)))
        da.fft(worker)
        db.fft(worker)
        da.mul_assign(worker, db)
        da.ifft(worker)
        got = bm.fr_from_mont(da.into_coeffs())
        assert got[: size] == naive and all(v == 0 for v in got[size:])
        da.free()
        db.free()


def test_pointwise_and_z(worker):
    n = 300            # pads to 512
    a, b = _rand(n, 1), _rand(n, 2)
    da = bm.EvaluationDomain.from_coeffs(worker, bm.fr_to_mont(a))
    db = bm.EvaluationDomain.from_coeffs(worker, bm.fr_to_mont(b))
    assert len(da) == 512 and da.exp == 9
    oa, ob = odomain.EvaluationDomain(F, a), odomain.EvaluationDomain(F, b)
    da.sub_assign(worker, db); oa.sub_assign(ob)
    da.mul_assign(worker, db); oa.mul_assign(ob)
    da.divide_by_z_on_coset(worker); oa.divide_by_z_on_coset()
    g = 0x1234567
    da.distribute_powers(worker, bm.fr_to_mont([g])[0]); oa.distribute_powers(g)
    assert bm.fr_from_mont(da.into_coeffs()) == oa.coeffs
    tau = 987654321
    assert bm.fr_from_mont(da.z(bm.fr_to_mont([tau])[0]).reshape(1, 4)) == [oa.z(tau)]
    short = bm.EvaluationDomain.from_coeffs(worker, bm.fr_to_mont(a[:100]))
    with pytest.raises(AssertionError):
        da.mul_assign(worker, short)
    for d in (da, db, short):
        d.free()


def test_h_coefficients(worker):
    """prover.rs:210-231 fused pipeline vs the step-by-step restatement"""
    from oracle import groth16 as og
    for n in (1, 5, 646, 1500):
        a, b, c = _rand(n, 11), _rand(n, 12), _rand(n, 13)
        got = bm.limbs_to_ints(bm.h_coefficients(worker, bm.fr_to_mont(a), bm.fr_to_mont(b), bm.fr_to_mont(c)))
        assert got == og.h_coefficients(F, a, b, c), n


def test_large_roundtrip_and_spot(worker):
    """2^20: ifft(fft(x)) == x and out[k] = sum_j a_j w^{jk} at random k (size-independent checks)"""
    logn = 20
    n = 1 << logn
    rs = np.random.RandomState(3)
    raw = rs.randint(0, 1 << 62, size=(n, 4), dtype=np.int64).astype(np.uint64)   # < 2^254 < q
    d = bm.EvaluationDomain.from_coeffs(worker, raw)
    d.fft(worker)
    out = d.into_coeffs()
    d.ifft(worker)
    assert np.array_equal(d.into_coeffs(), raw)
    d.free()
    # linear spot check against the definition at 3 output indices
    omega = pow(F.root_of_unity, 1 << (32 - logn), Q)
    a = bm.fr_from_mont(raw)
    for k in (0, 1, 777777):
        wk = pow(omega, k, Q)
        acc, w = 0, 1
        for v in a:
            acc = (acc + v * w) % Q
            w = w * wk % Q
        assert bm.fr_from_mont(out[k : k + 1]) == [acc]
