"""GPU parity: multiexp vs the oracle restatement of src/multiexp.rs.
`test_with_bls12` (multiexp.rs:283-327: naive sum == multiexp, FullDensity) is mirrored, and the
semantics the reference never unit-tests directly -- density maps, base offsets, EOF / identity
errors and their precedence (SURVEY 8a') -- are checked against `oracle.multiexp`."""
import random

import numpy as np
import pytest

import bellman_mpc_b200 as bm
from oracle import curves, fields
from oracle import multiexp as ome
from util import Q, decode, expected_from_dlogs, known_dlog_bases, rand_scalars

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["xyzz", "affine", "pairs", "affine_radix"])
def accumulate_kernel(request, monkeypatch, worker):
    """Every case runs through all three bucket-accumulation kernels: the XYZZ chain, the
    batched-affine tree (msm_affine.cuh; forced here, with 3 slices per job) and the rounds of pair
    additions (msm_pairs.cuh; forced: the automatic choice only takes them for multiexps that fill
    the GPU).  `affine_radix` adds the two-level partition sort (rs_* kernels) at every size the
    geometry allows, in place of the one-pass scatter the automatic choice keeps for small calls."""
    monkeypatch.setenv("BMPC_ACC_PAIRS", "1" if request.param == "pairs" else "0")
    monkeypatch.setenv("BMPC_ACC_AFFINE", "1" if request.param.startswith("affine") else "0")
    monkeypatch.setenv("BMPC_SORT_RADIX", "1" if request.param == "affine_radix" else "0")
    monkeypatch.setenv("BMPC_AFF_FORCE_G", "3")
    worker.reload_env()
    yield request.param
    monkeypatch.undo()
    worker.reload_env()


def _oracle_points(group, ks):
    G = curves.G1 if group == bm.G1 else curves.G2
    return [G.mul(G.gen, k) if k else None for k in ks]


def _run(worker, bases, start, density, scalars):
    return bm.multiexp(worker, (bases, start), density, bm.ints_to_limbs(scalars)).wait()


@pytest.mark.parametrize("group", [bm.G1, bm.G2])
@pytest.mark.parametrize("n", [1, 2, 31, 32, 200])
def test_with_bls12_small(worker, group, n):
    """naive sum over oracle points == GPU multiexp == oracle multiexp (FullDensity)"""
    G = curves.G1 if group == bm.G1 else curves.G2
    ks = rand_scalars(n, 10 + n)
    scalars = rand_scalars(n, 20 + n)
    pts = _oracle_points(group, ks)
    bases = bm.Bases.from_uncompressed(worker, group, b"".join(G.to_uncompressed(p) for p in pts))
    got = _run(worker, bases, 0, bm.FullDensity(), scalars)
    assert decode(group, got) == ome.naive(G, pts, scalars)
    assert decode(group, got) == ome.multiexp(G, pts, 0, ome.FullDensity(), scalars)
    # round trip of the resident copy (from_uncompressed -> Montgomery -> to_uncompressed)
    assert bases.read() == b"".join(G.to_uncompressed(p) for p in pts)
    bases.free()


@pytest.mark.parametrize("group,n", [(bm.G1, 1 << 10), (bm.G1, 1 << 14), (bm.G2, 1 << 10)])
@pytest.mark.parametrize("kind", ["uniform", "mixed", "small"])
def test_known_dlog(worker, group, n, kind):
    """bases with known discrete logs: expected = (sum k_i s_i) G, any size, no CPU MSM"""
    ks = rand_scalars(n, 2)
    scalars = rand_scalars(n, 1, kind)
    bases = known_dlog_bases(worker, group, ks)
    got = _run(worker, bases, 0, bm.FullDensity(), scalars)
    assert got == expected_from_dlogs(group, ks, scalars)
    bases.free()


@pytest.mark.parametrize("c", [2, 5, 9, 13, 16])
def test_window_independence(worker, c):
    """SURVEY 8a'/7: the result does not depend on the window size"""
    n = 3000
    ks, scalars = rand_scalars(n, 3), rand_scalars(n, 4, "mixed")
    bases = known_dlog_bases(worker, bm.G1, ks)
    worker.set_tuning(c, 0)
    try:
        assert _run(worker, bases, 0, bm.FullDensity(), scalars) == expected_from_dlogs(bm.G1, ks, scalars)
    finally:
        worker.set_tuning(0, 0)
    bases.free()


@pytest.mark.parametrize("group", [bm.G1, bm.G2])
def test_density_and_offset(worker, group):
    """SURVEY 8a'/1: k-th dense exponent consumes base start + k"""
    n, start = 777, 3
    rng = random.Random(9)
    bits = [rng.random() < 0.5 for _ in range(n)]
    ks = rand_scalars(start + sum(bits) + 5, 5)          # bases may be longer than needed
    scalars = rand_scalars(n, 6, "mixed")
    bases = known_dlog_bases(worker, group, ks)
    got = _run(worker, bases, start, bm.DensityTracker.from_bits(bits), scalars)
    assert got == expected_from_dlogs(group, ks, scalars, bits, start)
    # all-zero density: identity, nothing consumed even though start is past the end
    none = bm.DensityTracker.from_bits([False] * n)
    ident = _run(worker, bases, len(ks) + 10, none, scalars)
    assert decode(group, ident) is None
    bases.free()


def test_duplicate_and_opposite_bases(worker):
    """equal bases meet in one bucket (P + P) and opposite ones cancel (P + (-P))"""
    G = curves.G1
    n = 64
    k = 123456789
    ks = [k] * (n // 2) + [Q - k] * (n // 2)
    scalars = [5] * n
    bases = known_dlog_bases(worker, bm.G1, ks)
    got = _run(worker, bases, 0, bm.FullDensity(), scalars)
    assert decode(bm.G1, got) is None
    scalars2 = [5] * (n // 2) + [0] * (n // 2)
    got2 = _run(worker, bases, 0, bm.FullDensity(), scalars2)
    assert decode(bm.G1, got2) == G.mul(G.gen, 5 * k * (n // 2) % Q)
    bases.free()


def test_empty(worker):
    """SURVEY 8a'/8"""
    bases = known_dlog_bases(worker, bm.G1, [1, 2, 3])
    got = bm.multiexp(worker, (bases, 0), bm.FullDensity(), np.zeros((0, 4), dtype=np.uint64)).wait()
    assert decode(bm.G1, got) is None
    bases.free()


def test_density_length_assert(worker):
    """multiexp.rs:273-278"""
    bases = known_dlog_bases(worker, bm.G1, [1, 2, 3])
    with pytest.raises(AssertionError):
        bm.multiexp(worker, (bases, 0), bm.DensityTracker.from_bits([True, True]), bm.ints_to_limbs([1, 2, 3]))
    bases.free()


def _error_case(worker, pts, start, bits, scalars):
    """run GPU and oracle on the same case; both must agree on value or error class"""
    G = curves.G1
    bases = bm.Bases.from_uncompressed(worker, bm.G1, b"".join(G.to_uncompressed(p) for p in pts))
    dens_g = bm.FullDensity() if bits is None else bm.DensityTracker.from_bits(bits)
    dens_o = ome.FullDensity()
    if bits is not None:
        dens_o = ome.DensityTracker()
        dens_o.bv = list(bits)
    try:
        exp = ("ok", ome.multiexp(G, pts, start, dens_o, scalars))
    except ome.UnexpectedIdentity:
        exp = ("identity", None)
    except ome.UnexpectedEof:
        exp = ("eof", None)
    try:
        got = ("ok", decode(bm.G1, _run(worker, bases, start, dens_g, scalars)))
    except bm.UnexpectedIdentity:
        got = ("identity", None)
    except bm.UnexpectedEof:
        got = ("eof", None)
    bases.free()
    assert got == exp, (got[0], exp[0])
    return exp[0]


def test_error_semantics(worker):
    """SURVEY 8a'/3-5 (multiexp.rs:55-65,74-80,244-249)"""
    G = curves.G1
    rng = random.Random(77)
    n = 40                                   # reference window c = ceil(ln 40) = 4
    pts = [G.mul(G.gen, rng.randrange(1, Q)) for _ in range(n)]
    sc = [rng.randrange(Q) for _ in range(n)]
    # identity base under a zero scalar: fine
    p1 = list(pts); p1[7] = None
    s1 = list(sc); s1[7] = 0
    assert _error_case(worker, p1, 0, None, s1) == "ok"
    # identity under a non-zero scalar: UnexpectedIdentity
    assert _error_case(worker, p1, 0, None, sc) == "identity"
    # identity under density-0 position: fine (never consumed)
    bits = [True] * n; bits[7] = False
    p2 = list(pts); p2[n - 1] = None         # only 39 dense positions: base 39 is never consumed
    assert _error_case(worker, p2, 0, bits, sc) == "ok"
    # bases run out: EOF, also when the overrunning scalars are zero
    assert _error_case(worker, pts[:30], 0, None, sc) == "eof"
    s3 = list(sc)
    for i in range(30, n):
        s3[i] = 0
    assert _error_case(worker, pts[:30], 0, None, s3) == "eof"
    # start offset past the end with dense positions
    assert _error_case(worker, pts, n, None, sc) == "eof"
    # EOF + earlier identity whose top-window digit is zero -> EOF wins; non-zero -> identity wins
    p4 = list(pts[:30]); p4[3] = None
    s4 = list(sc); s4[3] = 5                 # small scalar: top window (bits 252..255) digit is 0
    assert _error_case(worker, p4, 0, None, s4) == "eof"
    s5 = list(sc); s5[3] = (1 << 253) + 9    # top-window digit non-zero
    assert _error_case(worker, p4, 0, None, s5) == "identity"
    # scalar == 1 on an identity base is consumed (window 0) -> identity error
    s6 = list(sc); s6[7] = 1
    assert _error_case(worker, p1, 0, None, s6) == "identity"


@pytest.mark.parametrize("world", [2, 3])
def test_error_precedence_across_shards(worker, world):
    """multiexp.rs:244-249 over a sharded multiexp: the shards run one after the other on
    this GPU through bmpc_multiexp_shard_dev, each against ITS slice of the bases; the OR of their
    raw flag words must give the status the oracle's multiexp reports for the whole vector.  The
    reference's window is c = ceil(ln 70) = 5 (top window = bits 250..254); a shard of 35 or 24
    exponents alone would take c = 4 (bits 252..255) and misjudge the digit 1 << 250."""
    import ctypes as C
    import torch
    from bellman_mpc_b200 import dist as bdist
    G = curves.G1
    rng = random.Random(91)
    n, nbases = 70, 60
    pts = [G.mul(G.gen, rng.randrange(1, Q)) for _ in range(nbases)]
    pts[3] = None                                                    # identity base on shard 0
    base_sc = [rng.randrange(Q) for _ in range(n)]
    lib = worker._lib
    for s3, want in (((1 << 250) + 9, "identity"), ((1 << 249) + 9, "eof"), (1, "eof"), (0, "eof")):
        sc = list(base_sc)
        sc[3] = s3
        try:
            ome.multiexp(G, pts, 0, ome.FullDensity(), sc)
            exp = "ok"
        except ome.UnexpectedIdentity:
            exp = "identity"
        except ome.UnexpectedEof:
            exp = "eof"
        assert exp == want
        acc = 0
        for r in range(world):
            lo, hi = bdist.shard_range(n, world, r)
            sl = pts[lo:min(hi, nbases)]                              # the rank's slice of the bases
            bases = bm.Bases.from_uncompressed(worker, bm.G1, b"".join(G.to_uncompressed(p) for p in sl)) if sl else \
                bm.Bases.from_uncompressed(worker, bm.G1, b"", 0)
            d_sc = torch.from_numpy(bm.ints_to_limbs(sc[lo:hi]).view(np.int64)).cuda()
            part = torch.zeros(int(lib.bmpc_partial_bytes(bm.G1)), dtype=torch.uint8, device="cuda")
            fl = C.c_uint32(0)
            rc = lib.bmpc_multiexp_shard_dev(worker.ctx, bases.handle, 0, d_sc.data_ptr(), hi - lo, None, 0, n,
                                             part.data_ptr(), C.byref(fl), None)
            assert rc == 0
            acc |= fl.value
            bases.free()
        got = {0: "ok", 1: "identity", 2: "eof"}[lib.bmpc_msm_flags_status(acc)]
        assert got == exp == {0: "ok", 1: "identity", 2: "eof"}[bdist.flags_status(acc)]


@pytest.mark.parametrize("group,world", [(bm.G1, 2), (bm.G1, 5), (bm.G2, 3)])
def test_shard_records_one_sync(worker, group, world):
    """The enqueue-only shard call and the one-synchronisation fold (bmpc_multiexp_shard_enqueue_dev /
    bmpc_fold_shard_records): `world` shards run back to back on this GPU, each leaving its record
    (XYZZ partial + raw flag word) in ONE buffer as an all-gather would; the fold must return the
    bytes of the whole multiexp and the OR of the flag words.  Also: an identity base consumed by a
    shard shows up in the folded flags, a rank with nothing to do contributes the identity."""
    import ctypes as C
    import torch
    from bellman_mpc_b200 import dist as bdist
    lib = worker._lib
    n = 1000
    ks, scalars = rand_scalars(n, 61), list(rand_scalars(n, 62, "mixed"))
    scalars[bdist.shard_range(n, world, world - 1)[0]] = 12345       # the poisoned base is consumed
    rb = int(lib.bmpc_shard_record_bytes(group))
    assert rb == int(lib.bmpc_partial_bytes(group)) + 16
    for poison in (False, True):
        recs = torch.full((world * rb,), 0xAB, dtype=torch.uint8, device="cuda")   # stale bytes must not leak
        keep = []
        for r in range(world):
            lo, hi = bdist.shard_range(n, world, r)
            sl_ks = list(ks[lo:hi])
            if poison and r == world - 1:
                sl_ks[0] = 0                                          # identity base on the last shard
            bases = known_dlog_bases(worker, group, sl_ks)
            d_sc = torch.from_numpy(bm.ints_to_limbs(scalars[lo:hi]).view(np.int64)).cuda()
            rc = lib.bmpc_multiexp_shard_enqueue_dev(worker.ctx, bases.handle, 0, d_sc.data_ptr(), hi - lo, None, 0, n,
                                                     recs.data_ptr() + r * rb, None)
            assert rc == 0, lib.bmpc_last_error(worker.ctx)
            keep.append((bases, d_sc))
        out = np.zeros(96 if group == bm.G1 else 192, dtype=np.uint8)
        fl = C.c_uint32(0xffffffff)
        rc = lib.bmpc_fold_shard_records(worker.ctx, group, recs.data_ptr(), world, rb, out.ctypes.data_as(C.c_void_p),
                                         C.byref(fl), None)
        assert rc == 0, lib.bmpc_last_error(worker.ctx)
        if poison:
            assert fl.value & 2 and lib.bmpc_msm_flags_status(fl.value) == bm._lib.ERR_UNEXPECTED_IDENTITY
        else:
            assert fl.value == 0
            assert out.tobytes() == expected_from_dlogs(group, ks, scalars)
        for bases, _ in keep:
            bases.free()
    # n == 0 on a rank: identity partial, zero flags
    recs = torch.full((rb,), 0xCD, dtype=torch.uint8, device="cuda")
    bases = known_dlog_bases(worker, group, ks[:4])
    assert lib.bmpc_multiexp_shard_enqueue_dev(worker.ctx, bases.handle, 0, None, 0, None, 0, n, recs.data_ptr(), None) == 0
    out = np.zeros(96 if group == bm.G1 else 192, dtype=np.uint8)
    fl = C.c_uint32(7)
    assert lib.bmpc_fold_shard_records(worker.ctx, group, recs.data_ptr(), 1, rb, out.ctypes.data_as(C.c_void_p),
                                       C.byref(fl), None) == 0
    assert fl.value == 0 and decode(group, out.tobytes()) is None
    bases.free()


@pytest.mark.parametrize("group,n,c", [(bm.G1, 5000, 0), (bm.G1, 3000, 9), (bm.G2, 1500, 0), (bm.G1, 1 << 16, 0)])
def test_precomputed_tables(worker, group, n, c):
    """window tables 2^(cw) P_i (one bucket set, no doubling fold) give the same bytes, also with
    density maps, offsets and the error semantics"""
    ks = rand_scalars(n + 7, 50)
    scalars = rand_scalars(n, 51, "mixed")
    bases = known_dlog_bases(worker, group, ks).precompute(c)
    assert _run(worker, bases, 0, bm.FullDensity(), scalars) == expected_from_dlogs(group, ks, scalars)
    rng = random.Random(52)
    bits = [rng.random() < 0.5 for _ in range(n)]
    got = _run(worker, bases, 5, bm.DensityTracker.from_bits(bits), scalars)
    assert got == expected_from_dlogs(group, ks, scalars, bits, 5)
    with pytest.raises(bm.UnexpectedEof):
        _run(worker, bases, 8, bm.FullDensity(), scalars)
    assert bases.read(0, 3) == known_dlog_bases(worker, group, ks[:3]).read()
    bases.free()


@pytest.mark.parametrize("group,c", [(bm.G1, 12), (bm.G1, 10), (bm.G2, 12), (bm.G1, 13)])
@pytest.mark.parametrize("n", [1, 16, 300])
def test_precomputed_tables_short_multiexp(worker, group, c, n):
    """A multiexp over a few points of a table-registered vector (create_proof's input multiexps:
    16 exponents against a 2^22-point query vector) splits the table window into sub-windows with
    their own small bucket sets (msm_make_plan); c = 13 is prime and keeps the one-set geometry.
    Same bytes, incl. an offset into the vector and a density map."""
    ks = rand_scalars(4000, 60)
    bases = known_dlog_bases(worker, group, ks).precompute(c)
    scalars = rand_scalars(n, 61 + n, "mixed" if n > 1 else "uniform")
    assert _run(worker, bases, 0, bm.FullDensity(), scalars) == expected_from_dlogs(group, ks, scalars)
    rng = random.Random(62)
    bits = [rng.random() < 0.6 for _ in range(n)]
    got = _run(worker, bases, 3000, bm.DensityTracker.from_bits(bits), scalars)
    assert got == expected_from_dlogs(group, ks, scalars, bits, 3000)
    bases.free()


def test_precomputed_identity_base(worker):
    G = curves.G1
    rng = random.Random(53)
    n = 600
    pts = [G.mul(G.gen, rng.randrange(1, Q)) for _ in range(n)]
    pts[17] = None
    sc = [rng.randrange(Q) for _ in range(n)]
    bases = bm.Bases.from_uncompressed(worker, bm.G1, b"".join(G.to_uncompressed(p) for p in pts)).precompute(8)
    with pytest.raises(bm.UnexpectedIdentity):
        _run(worker, bases, 0, bm.FullDensity(), sc)
    sc[17] = 0
    assert decode(bm.G1, _run(worker, bases, 0, bm.FullDensity(), sc)) == ome.naive(G, pts[:17] + pts[18:], sc[:17] + sc[18:])
    bases.free()


def test_batch_scalar_mul(worker):
    """mpc.rs:647-706: per-element and same-scalar batch multiplication, G1 and G2"""
    for group in (bm.G1, bm.G2):
        G = curves.G1 if group == bm.G1 else curves.G2
        ks = rand_scalars(50, 31)
        mult = rand_scalars(50, 32)
        mult[3] = 0
        bases = known_dlog_bases(worker, group, ks)
        per = bases.scalar_mul(bm.ints_to_limbs(mult), per_element=True)
        exp = b"".join(G.to_uncompressed(G.mul(G.gen, k * m % Q)) for k, m in zip(ks, mult))
        assert per.read() == exp
        same = bases.scalar_mul(bm.ints_to_limbs([mult[0]]), per_element=False)
        exp2 = b"".join(G.to_uncompressed(G.mul(G.gen, k * mult[0] % Q)) for k in ks)
        assert same.read() == exp2
        for b in (bases, per, same):
            b.free()


def test_large_known_dlog(worker):
    """config #2: G1 multiexp 2^20, uniform scalars, known-dlog bases, FullDensity"""
    n = 1 << 20
    rs = np.random.RandomState(1)
    ks_l = rs.randint(0, 1 << 62, size=(n, 4), dtype=np.int64).astype(np.uint64)
    sc_l = rs.randint(0, 1 << 62, size=(n, 4), dtype=np.int64).astype(np.uint64)
    G = curves.G1
    bases = bm.Bases.fixed_base_mul(worker, bm.G1, G.to_uncompressed(G.gen), ks_l)
    got = bm.multiexp(worker, (bases, 0), bm.FullDensity(), sc_l).wait()
    ks, sc = bm.limbs_to_ints(ks_l), bm.limbs_to_ints(sc_l)
    dot = sum(k * s for k, s in zip(ks, sc)) % Q
    assert got == G.to_uncompressed(G.mul(G.gen, dot))
    bases.free()


@pytest.mark.parametrize("tables,log_n", [(False, 18), (True, 18), (True, 20)])
def test_hot_buckets_large(worker, tables, log_n):
    """boolean-heavy witness shape at 2^18 (SURVEY 8d/4 secondary profile): 35 % zeros, 35 % ones,
    10 % small values, 20 % uniform -- a few buckets receive a large share of the points (task
    splitting, heavy-bucket combine, warp-aggregated atomics), also at 2^20; checked against
    sum k_i s_i"""
    from oracle import cref
    n = 1 << log_n
    rs = np.random.RandomState(77)
    ks = rs.randint(0, 1 << 62, size=(n, 4), dtype=np.int64).astype(np.uint64)
    sc = rs.randint(0, 1 << 62, size=(n, 4), dtype=np.int64).astype(np.uint64)
    kind = rs.randint(0, 100, size=n)
    sc[kind < 35] = 0
    ones = (kind >= 35) & (kind < 70)
    sc[ones] = 0
    sc[ones, 0] = 1
    small = (kind >= 70) & (kind < 80)
    sc[small, 1:] = 0
    sc[small, 0] &= np.uint64(0xFFFF)
    G = curves.G1
    bases = bm.Bases.fixed_base_mul(worker, bm.G1, G.to_uncompressed(G.gen), ks)
    if tables:
        bases.precompute()
    got = bm.multiexp(worker, (bases, 0), bm.FullDensity(), sc).wait()
    assert got == cref.g1_generator_mul(cref.fr_dot(ks, sc))
    bases.free()


def test_list_mul_matrix(worker):
    """mpc.rs:416-457: rows of a sparse matrix times a point list, G1 and G2 in one call; the
    reference stops at the first empty row and leaves the rest of the result at the identity"""
    from oracle import mpc as ompc
    rng = random.Random(77)
    n = 70
    ks = rand_scalars(n, 71)
    ks[5] = 0                                        # an identity element in the list
    b1, b2 = known_dlog_bases(worker, bm.G1, ks), known_dlog_bases(worker, bm.G2, ks)
    matrix = []
    for i in range(64):
        k = 0 if i == 50 else rng.randrange(1, 6)
        matrix.append([(rng.choice([0, 1, 2, Q - 1, rng.randrange(Q), rng.randrange(1 << 40)]), rng.randrange(n))
                       for _ in range(k)])
    matrix[3] = [(rng.randrange(Q), 5), (0, 7)]      # a row summing to the identity
    matrix[4] = [(7, 9), (Q - 7, 9)]                 # P - P
    matrix[6] = [(rng.randrange(Q), 11)] * 2         # the same entry twice (doubling inside the row)
    r1, r2 = bm.list_mul_matrix(b1, b2, matrix)
    assert len(r1) == n and len(r2) == n
    for G, got in ((curves.G1, r1.read()), (curves.G2, r2.read())):
        want = b""
        for i in range(n):
            dot = sum(cf * ks[idx] for cf, idx in matrix[i]) % Q if i < 50 else 0
            want += G.to_uncompressed(G.mul(G.gen, dot))
        assert got == want
    # the restatement of the reference on the same inputs (small case: big-integer curve arithmetic)
    G = curves.G1
    lst = [G.mul(G.gen, k) for k in ks[:12]]
    small = [[(cf, idx % 12) for cf, idx in row] for row in matrix[:8]]
    s1 = bm.Bases.from_uncompressed(worker, bm.G1, b"".join(G.to_uncompressed(p) for p in lst))
    s2 = known_dlog_bases(worker, bm.G2, ks[:12])
    o1, o2 = bm.list_mul_matrix(s1, s2, small)
    assert o1.read() == b"".join(G.to_uncompressed(p) for p in ompc.list_mul_matrix(G, lst, small))
    # empty matrix -> all identity; index panics of the reference -> AssertionError
    e1, e2 = bm.list_mul_matrix(s1, s2, [])
    assert e1.read() == G.to_uncompressed(None) * 12
    with pytest.raises(AssertionError):
        bm.list_mul_matrix(s1, s2, [[(1, 12)]])
    with pytest.raises(AssertionError):
        bm.list_mul_matrix(s1, s2, [[(1, 0)]] * 13)
    # a matrix TALLER than the list whose first empty row comes before row list.len() is fine in the
    # reference (the loop breaks there, mpc.rs:432-434) -- and here
    tall = [[(3, 1)], [(5, 2)], []] + [[(1, 0)]] * 20
    t1, t2 = bm.list_mul_matrix(s1, s2, tall)
    assert t1.read() == b"".join(G.to_uncompressed(p) for p in ompc.list_mul_matrix(G, lst, tall))
    t1.free(); t2.free()
    for b in (b1, b2, r1, r2, s1, s2, o1, o2, e1, e2):
        b.free()
