"""GPU parity for the two host-facing forms added on top of the single-device calls:

* the asynchronous Waiter (multicore.rs:33-118; prover.rs:233-307 keeps eight multiexps in flight):
  bmpc_multiexp_async / bmpc_waiter_wait give the same bytes and statuses as bmpc_multiexp with any
  number in flight;
* the multi-device context (SURVEY 8b `bmpc_ctx_create(devices, n)`): multiexp and create_proof as ONE
  call over N devices, here with the same GPU named N times so the whole plan -- even base split, position
  cuts through the density map, per-device scalars and density slices, peer gather, fold, flag
  precedence -- runs on the one-GPU test box.  Results must equal the single-device ones byte for
  byte and the oracle's statuses."""
import ctypes as C
import random

import numpy as np
import pytest

import bellman_mpc_b200 as bm
from oracle import curves, fields
from oracle import multiexp as ome
from util import Q, decode, expected_from_dlogs, known_dlog_bases, rand_scalars

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["one-gpu-three-times", "all-gpus"])
def multi3(request):
    """the same GPU named three times (runs the whole plan on a one-GPU box), and, where the box has
    several, all of them: the partial sums then really cross devices (cudaMemcpyPeerAsync)"""
    if request.param == "all-gpus":
        import torch
        nd = torch.cuda.device_count()
        if nd < 2:
            pytest.skip("one GPU only")
        w = bm.MultiWorker(list(range(min(nd, 8))))
    else:
        w = bm.MultiWorker([0, 0, 0])
    yield w
    w.close()


# ------------------------------------------------------------------------------ async waiter
def test_eight_multiexps_in_flight(worker):
    """the shape of prover.rs:233-307: all multiexps are issued before the first wait()"""
    cases = []
    for j in range(8):
        grp = bm.G2 if j in (4, 5) else bm.G1
        n = [1, 700, 33, 4096, 5, 900, 2047, 1500][j]
        ks, sc = rand_scalars(n + 4, 300 + j), rand_scalars(n, 320 + j, "mixed")
        bits = None
        if j % 2:
            rng = random.Random(j)
            bits = [rng.random() < 0.6 for _ in range(n)]
        cases.append((grp, ks, sc, bits, known_dlog_bases(worker, grp, ks)))
    waiters = []
    for grp, ks, sc, bits, bases in cases:
        dens = bm.FullDensity() if bits is None else bm.DensityTracker.from_bits(bits)
        waiters.append(bm.multiexp(worker, (bases, 2 if bits else 0), dens, bm.ints_to_limbs(sc)))
    for (grp, ks, sc, bits, bases), wt in reversed(list(zip(cases, waiters))):     # any wait order
        assert wt.wait() == expected_from_dlogs(grp, ks, sc, bits, 2 if bits else 0)
    for c in cases:
        c[4].free()


def test_waiter_statuses(worker):
    G = curves.G1
    pts = [G.mul(G.gen, k) for k in range(1, 31)]
    p_id = list(pts)
    p_id[4] = None
    b_ok = bm.Bases.from_uncompressed(worker, bm.G1, b"".join(G.to_uncompressed(p) for p in pts))
    b_id = bm.Bases.from_uncompressed(worker, bm.G1, b"".join(G.to_uncompressed(p) for p in p_id))
    sc = rand_scalars(30, 5)
    w_eof = bm.multiexp(worker, (b_ok, 1), bm.FullDensity(), bm.ints_to_limbs(sc))       # 30 exponents, 29 bases
    w_id = bm.multiexp(worker, (b_id, 0), bm.FullDensity(), bm.ints_to_limbs(sc))
    w_ok = bm.multiexp(worker, (b_ok, 0), bm.FullDensity(), bm.ints_to_limbs(sc))
    w_empty = bm.multiexp(worker, (b_ok, 0), bm.FullDensity(), np.zeros((0, 4), dtype=np.uint64))
    assert decode(bm.G1, w_ok.wait()) == ome.multiexp(G, pts, 0, ome.FullDensity(), sc)
    with pytest.raises(bm.UnexpectedIdentity):
        w_id.wait()
    with pytest.raises(bm.UnexpectedEof):
        w_eof.wait()
    assert decode(bm.G1, w_empty.wait()) is None
    # raw C ABI: same bytes as the blocking call
    lib = worker._lib
    limbs = bm.ints_to_limbs(sc)
    out_a, out_b = np.zeros(96, dtype=np.uint8), np.zeros(96, dtype=np.uint8)
    h = C.c_void_p()
    assert lib.bmpc_multiexp_async(worker.ctx, b_ok.handle, 0, limbs.ctypes.data_as(C.c_void_p), 30, None, 0, C.byref(h)) == 0
    assert lib.bmpc_multiexp(worker.ctx, b_ok.handle, 0, limbs.ctypes.data_as(C.c_void_p), 30, None, 0,
                             out_b.ctypes.data_as(C.c_void_p)) == 0
    assert lib.bmpc_waiter_wait(h, out_a.ctypes.data_as(C.c_void_p)) == 0
    assert out_a.tobytes() == out_b.tobytes()
    b_ok.free()
    b_id.free()


# ------------------------------------------------------------------------------ multi-device
def _multi_bases(multi, group, ks):
    """known-dlog bases made on one device, then split over the devices of `multi`"""
    w = bm.Worker(0)
    b = known_dlog_bases(w, group, ks)
    raw = b.read()
    b.free()
    w.close()
    return bm.MultiBases.from_uncompressed(multi, group, raw)


@pytest.mark.parametrize("group,n", [(bm.G1, 5000), (bm.G2, 1000), (bm.G1, 2), (bm.G1, 1 << 16)])
def test_multi_multiexp_full_density(multi3, group, n):
    ks, sc = rand_scalars(n, 400 + n % 97), rand_scalars(n, 401, "mixed")
    mb = _multi_bases(multi3, group, ks)
    assert len(mb) == n
    got = bm.multiexp(multi3, (mb, 0), bm.FullDensity(), bm.ints_to_limbs(sc)).wait()
    assert got == expected_from_dlogs(group, ks, sc)
    if n >= 5000:
        mb.precompute()
        assert bm.multiexp(multi3, (mb, 0), bm.FullDensity(), bm.ints_to_limbs(sc)).wait() == got
    mb.free()


@pytest.mark.parametrize("start,p_dense", [(0, 0.5), (3, 0.5), (700, 0.9), (3, 0.02), (1999, 0.5)])
def test_multi_multiexp_density_and_offset(multi3, start, p_dense):
    """the k-th dense exponent consumes base start + k whichever device holds it; cuts fall at
    arbitrary bit positions of the density words"""
    n = 2500
    rng = random.Random(int(start * 10 + p_dense * 100))
    bits = [rng.random() < p_dense for _ in range(n)]
    ks = rand_scalars(start + sum(bits) + 11, 410)
    sc = rand_scalars(n, 411, "mixed")
    mb = _multi_bases(multi3, bm.G1, ks)
    got = bm.multiexp(multi3, (mb, start), bm.DensityTracker.from_bits(bits), bm.ints_to_limbs(sc)).wait()
    assert got == expected_from_dlogs(bm.G1, ks, sc, bits, start)
    mb.free()


def test_multi_error_precedence(multi3):
    """multiexp.rs:244-249 through the one-call form: an identity base on the first device whose digit in
    the reference's top window (c = ceil(ln 70) = 5: bits 250..254) is non-zero beats the overrun the
    last device sees; otherwise EOF; statuses equal to the oracle's on the whole vector."""
    G = curves.G1
    rng = random.Random(91)
    n, nbases = 70, 60
    pts = [G.mul(G.gen, rng.randrange(1, Q)) for _ in range(nbases)]
    pts[3] = None
    base_sc = [rng.randrange(Q) for _ in range(n)]
    mb = bm.MultiBases.from_uncompressed(multi3, bm.G1, b"".join(G.to_uncompressed(p) for p in pts))
    for s3, want in (((1 << 250) + 9, bm.UnexpectedIdentity), ((1 << 249) + 9, bm.UnexpectedEof),
                     (1, bm.UnexpectedEof), (0, bm.UnexpectedEof)):
        sc = list(base_sc)
        sc[3] = s3
        try:
            ome.multiexp(G, pts, 0, ome.FullDensity(), sc)
            exp = None
        except ome.UnexpectedIdentity:
            exp = bm.UnexpectedIdentity
        except ome.UnexpectedEof:
            exp = bm.UnexpectedEof
        assert exp is want
        with pytest.raises(want):
            bm.multiexp(multi3, (mb, 0), bm.FullDensity(), bm.ints_to_limbs(sc)).wait()
    # no overrun: the identity alone decides
    sc = list(base_sc[:nbases])
    with pytest.raises(bm.UnexpectedIdentity):
        bm.multiexp(multi3, (mb, 0), bm.FullDensity(), bm.ints_to_limbs(sc)).wait()
    sc[3] = 0
    got = bm.multiexp(multi3, (mb, 0), bm.FullDensity(), bm.ints_to_limbs(sc)).wait()
    assert decode(bm.G1, got) == ome.multiexp(G, pts, 0, ome.FullDensity(), sc)
    with pytest.raises(AssertionError):                                # multiexp.rs:273-278
        bm.multiexp(multi3, (mb, 0), bm.DensityTracker.from_bits([True] * 5), bm.ints_to_limbs(sc))
    mb.free()


@pytest.mark.parametrize("log_m,profile", [(5, "uniform"), (11, "uniform"), (12, "boolean")])
def test_multi_create_proof(worker, multi3, log_m, profile):
    """create_proof over three devices in one call == the single-device proof == the known-dlog
    expectation.  The CRS is split evenly by base index with no knowledge of the circuit; the
    position cuts of the six density-mapped multiexps come from the witness' density maps per call."""
    import bench_prove
    wl = bench_prove.Workload(worker, log_m, seed=60 + log_m, profile=profile)
    expect = wl.prove()
    assert expect == wl.expected_proof()
    split = lambda b: bm.MultiBases.from_uncompressed(multi3, b.group, b.read())
    p = wl.params
    mp = bm.MultiParameters(multi3, split(wl.h), split(wl.l), split(wl.qa), split(wl.qb1), split(wl.qb2),
                            p.alpha_g1, p.beta_g1, p.beta_g2, p.delta_g1, p.delta_g2)
    assert bm.create_proof(wl.assignment, mp, wl.r, wl.s) == expect
    # subversion check (prover.rs:309-313) and a truncated query vector (EOF) through the same call
    mp_bad = bm.MultiParameters(multi3, mp.h, mp.l, mp.a, mp.b_g1, mp.b_g2, p.alpha_g1, p.beta_g1, p.beta_g2,
                                curves.G1.to_uncompressed(None), p.delta_g2)
    with pytest.raises(bm.UnexpectedIdentity):
        bm.create_proof(wl.assignment, mp_bad, wl.r, wl.s)
    short_l = bm.MultiBases.from_uncompressed(multi3, bm.G1, wl.l.read()[:-96])
    mp_short = bm.MultiParameters(multi3, mp.h, short_l, mp.a, mp.b_g1, mp.b_g2, p.alpha_g1, p.beta_g1, p.beta_g2,
                                  p.delta_g1, p.delta_g2)
    with pytest.raises(bm.UnexpectedEof):
        bm.create_proof(wl.assignment, mp_short, wl.r, wl.s)
    short_l.free()
    mp.free()
    wl.free()


# ------------------------------------------------------------------------------ ceremony check
def test_folded_contribution_check(worker):
    """groth16/mpc.rs:1091-1124 batched by a random linear combination: the GPU folds the contributed
    vector and the stored one (two multiexps in flight), the oracle's pairing closes the check.  Accept /
    reject agree with the reference's per-element loop on a small case; a 2^12 case is checked through
    known discrete logarithms."""
    from oracle import groth16 as og
    from oracle import mpc as ompc
    E, G1, G2 = og.BLS12, curves.G1, curves.G2
    rng = random.Random(31)
    delta = rng.randrange(1, Q)
    d2 = G2.mul(G2.gen, delta)
    dinv = pow(delta, -1, Q)
    for n, corrupt in ((5, None), (5, 3), (1 << 12, None), (1 << 12, 77)):
        ms = [rng.randrange(1, Q) for _ in range(n)]
        news = [m * dinv % Q for m in ms]
        if corrupt is not None:
            news[corrupt] = (news[corrupt] + 1) % Q
        matrixed, new = known_dlog_bases(worker, bm.G1, ms), known_dlog_bases(worker, bm.G1, news)
        rho = [rng.randrange(1 << 128) for _ in range(n)]
        f_new, f_old = bm.fold_vectors(worker, [new, matrixed], bm.ints_to_limbs(rho))
        assert f_new == expected_from_dlogs(bm.G1, news, rho) and f_old == expected_from_dlogs(bm.G1, ms, rho)
        ok = ompc.verify_vector_folded(E, decode(bm.G1, f_new), d2, decode(bm.G1, f_old))
        assert ok == (corrupt is None)
        if n <= 8:      # the reference's loop: two pairings per element
            pts_new = [G1.mul(G1.gen, k) for k in news]
            pts_old = [G1.mul(G1.gen, k) for k in ms]
            assert ompc.verify_vector(E, pts_new, d2, pts_old) == ok
        matrixed.free()
        new.free()


def test_multi_corner_splits(multi3):
    """fewer bases than devices, an offset that skips whole devices, all-zero densities, and a base
    vector that is longer than any exponent vector uses"""
    nd = len(multi3)
    # two bases over >= 3 devices: at least one device holds nothing
    ks = rand_scalars(2, 430)
    mb = _multi_bases(multi3, bm.G1, ks)
    sc = rand_scalars(2, 431)
    assert bm.multiexp(multi3, (mb, 0), bm.FullDensity(), bm.ints_to_limbs(sc)).wait() == expected_from_dlogs(bm.G1, ks, sc)
    assert bm.multiexp(multi3, (mb, 1), bm.FullDensity(), bm.ints_to_limbs(sc[:1])).wait() == \
        expected_from_dlogs(bm.G1, ks[1:], sc[:1])
    with pytest.raises(bm.UnexpectedEof):
        bm.multiexp(multi3, (mb, 1), bm.FullDensity(), bm.ints_to_limbs(sc)).wait()
    mb.free()
    # the offset starts on the last device; the devices before it get nothing to do
    n = 64 * nd
    ks = rand_scalars(n, 432)
    mb = _multi_bases(multi3, bm.G1, ks)
    start = n - 20
    sc = rand_scalars(20, 433, "mixed")
    got = bm.multiexp(multi3, (mb, start), bm.FullDensity(), bm.ints_to_limbs(sc)).wait()
    assert got == expected_from_dlogs(bm.G1, ks[start:], sc)
    # nothing dense: identity, whatever the offset
    none = bm.DensityTracker.from_bits([False] * 300)
    got = bm.multiexp(multi3, (mb, n + 5), none, bm.ints_to_limbs(rand_scalars(300, 434))).wait()
    assert decode(bm.G1, got) is None
    # a sparse map whose dense positions all map into the first device's slice
    bits = [i % 50 == 0 for i in range(500)]
    sc = rand_scalars(500, 435)
    got = bm.multiexp(multi3, (mb, 0), bm.DensityTracker.from_bits(bits), bm.ints_to_limbs(sc)).wait()
    assert got == expected_from_dlogs(bm.G1, ks, sc, bits, 0)
    mb.free()
