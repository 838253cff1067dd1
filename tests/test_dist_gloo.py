"""world_size-2 gloo test of the sharded-multiexp host logic: slicing, base-offset
prefix popcount, all-gather of partials + statuses, fold.  The per-rank partial is produced by the
oracle over the reference's DummyEngine group standing in for the device (the reference's own
"fake backend" technique, groth16/tests/dummy_engine.rs) -- no GPU needed."""
import os
import random

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from bellman_mpc_b200 import _lib
from bellman_mpc_b200 import dist as bdist
from oracle import curves, fields
from oracle import multiexp as ome

P = fields.DummyFr.p


def _case(seed, n=203, nbases=None, ident=None):
    """ident = (position, exponent): the base that dense position consumes becomes the identity"""
    rng = random.Random(seed)
    bits = [rng.random() < 0.6 for _ in range(n)]
    if nbases is None:
        nbases = 3 + sum(bits) + 2
    bases = [rng.randrange(1, P) for _ in range(nbases)]
    exps = [rng.choice([0, 1, rng.randrange(P)]) for _ in range(n)]
    if ident is not None:
        pos, e = ident
        bits[pos] = True
        exps[pos] = e
        bases[3 + sum(bits[:pos])] = 0           # Dummy group: 0 is the identity
    return bases, exps, bits


def _words(bits):
    w = [0] * max(1, (len(bits) + 63) // 64)
    for i, b in enumerate(bits):
        if b:
            w[i // 64] |= 1 << (i % 64)
    return np.array(w, dtype=np.uint64)


def _worker(rank, world, port, seed, nbases, use_density, q, ident=None):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    G = curves.Dummy
    bases, exps, bits = _case(seed, nbases=nbases, ident=ident)
    words = _words(bits) if use_density else None

    def partial_fn(lo, hi, first_base, dens_slice, n_total):
        # what bmpc_multiexp_shard_dev hands back: the slice's partial sum + its raw flag word
        dbits = [True] * (hi - lo)
        if dens_slice is not None:
            dbits = [bool((int(dens_slice[i // 64]) >> (i % 64)) & 1) for i in range(hi - lo)]
        flags = ome.shard_flags(G, bases, first_base, dbits, exps[lo:hi], n_total, num_bits=16)
        if flags:
            return _lib.OK, flags, bytes(4)
        d = ome.FullDensity()
        if dens_slice is not None:
            d = ome.DensityTracker()
            d.bv = dbits
        v = ome.multiexp(G, bases, first_base, d, exps[lo:hi], num_bits=16)
        return _lib.OK, 0, int(v).to_bytes(4, "little")

    fold = lambda parts: sum(int.from_bytes(p, "little") for p in parts) % P
    st, res = bdist.sharded_multiexp(partial_fn, fold, len(exps), words, 3 if use_density else 0)
    q.put((rank, st, res))
    dist.destroy_process_group()


def _run(seed, nbases, use_density, port, ident=None):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, seed, nbases, use_density, q, ident)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    return sorted(res)


@pytest.mark.parametrize("use_density", [False, True])
def test_sharded_matches_single(use_density):
    seed = 5
    nb = None if use_density else 203 + 5
    bases, exps, bits = _case(seed, nbases=nb)
    d = ome.FullDensity()
    if use_density:
        d = ome.DensityTracker()
        d.bv = bits
    nb = len(bases)
    expect = ome.multiexp(curves.Dummy, bases, 3 if use_density else 0, d, exps, num_bits=16)
    res = _run(seed, nb, use_density, 29511 + int(use_density))
    assert [r[1] for r in res] == [_lib.OK, _lib.OK]
    assert res[0][2] == res[1][2] == expect


def test_sharded_eof_propagates():
    res = _run(6, 20, True, 29517)            # far too few bases: the upper rank overruns
    assert all(r[1] == _lib.ERR_UNEXPECTED_EOF for r in res)


def _whole_status(seed, nbases, ident):
    bases, exps, bits = _case(seed, nbases=nbases, ident=ident)
    d = ome.DensityTracker()
    d.bv = bits
    try:
        ome.multiexp(curves.Dummy, bases, 3, d, exps, num_bits=16)
        return _lib.OK
    except ome.UnexpectedEof:
        return _lib.ERR_UNEXPECTED_EOF
    except ome.UnexpectedIdentity:
        return _lib.ERR_UNEXPECTED_IDENTITY


@pytest.mark.parametrize("exp,port", [(0x3abc, 29521), (0x0abc, 29523), (1, 29525), (0, 29527)])
def test_identity_on_lower_rank_vs_overrun_on_upper_rank(exp, port):
    """multiexp.rs:244-249: the error reported is the first one, in scan order, of the HIGHEST failing
    window.  Rank 1 overruns the bases (fails every window); rank 0 holds an identity base whose
    exponent has a non-zero digit in the reference's top window (0x3abc, window 12..17 of c = 6 for
    n = 203) -> UnexpectedIdentity; a zero top digit (0x0abc), exponent 1 (consumed in window 0 only)
    -> UnexpectedEof; exponent 0 never looks at the base -> UnexpectedEof."""
    ident = (5, exp)
    nbases = 60                                   # rank 0's ~61 dense positions fit, rank 1 overruns
    expect = _whole_status(7, nbases, ident)
    assert expect == (_lib.ERR_UNEXPECTED_IDENTITY if exp == 0x3abc else _lib.ERR_UNEXPECTED_EOF)
    res = _run(7, nbases, True, port, ident)
    assert [r[1] for r in res] == [expect, expect]


def test_shard_flags_reproduce_the_whole_multiexp_status():
    """OR of the shards' flag words -> status == the oracle's multiexp over the whole vector, for
    random placements of identity bases and overruns, worlds 1..4 (CPU, no processes)."""
    G = curves.Dummy
    rng = random.Random(11)
    for trial in range(300):
        n = rng.randrange(1, 120)
        bits = [rng.random() < 0.7 for _ in range(n)]
        nb = max(1, sum(bits) + rng.choice([-7, -1, 0, 3]))
        bases = [rng.choice([0, rng.randrange(1, P)]) if rng.random() < 0.1 else rng.randrange(1, P) for _ in range(nb)]
        exps = [rng.choice([0, 1, rng.randrange(P), rng.randrange(1 << 10)]) for _ in range(n)]
        d = ome.DensityTracker()
        d.bv = bits
        try:
            ome.multiexp(G, bases, 0, d, exps, num_bits=16)
            expect = None
        except ome.UnexpectedEof:
            expect = "eof"
        except ome.UnexpectedIdentity:
            expect = "identity"
        for world in (1, 2, 3, 4):
            acc = 0
            for r in range(world):
                lo, hi = bdist.shard_range(n, world, r)
                acc |= ome.shard_flags(G, bases, sum(bits[:lo]), bits[lo:hi], exps[lo:hi], n, num_bits=16)
            assert ome.flags_status(acc) == expect, (trial, world)
            want = {None: _lib.OK, "eof": _lib.ERR_UNEXPECTED_EOF, "identity": _lib.ERR_UNEXPECTED_IDENTITY}[expect]
            assert bdist.flags_status(acc) == want


def test_slicing_helpers():
    assert [bdist.shard_range(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    bits = [(i * 7) % 3 == 0 for i in range(200)]
    w = _words(bits)
    for lo in (0, 1, 63, 64, 65, 130, 200):
        assert bdist.dense_before(w, lo) == sum(bits[:lo])
    s = bdist.slice_density(w, 70, 150)
    assert [bool((int(s[i // 64]) >> (i % 64)) & 1) for i in range(80)] == bits[70:150]
    assert bdist.dense_before(None, 17) == 17 and bdist.slice_density(None, 0, 5) is None


def test_proof_shard_plan_covers_the_reference_mapping():
    """ProofShardPlan: per multiexp, the ranks' (exponent range, first base) pairs tile the
    reference's single scan (k-th dense exponent consumes base start + k, multiexp.rs:191-222) and
    the slices of the query vectors are contiguous, disjoint and complete."""
    import numpy as np

    from bellman_mpc_b200 import dist as bdist
    rs = np.random.RandomState(3)
    for ni, na, world in ((2, 645, 2), (16, 1008, 3), (16, 4080, 8), (3, 50, 4)):
        m = 1
        while m < ni + na:
            m *= 2
        bits = lambda n, p: rs.random_sample(n) < p
        a_bits, bi_bits, ba_bits = bits(na, 0.75), bits(ni, 0.5), bits(na, 0.5)
        pack = lambda b: np.packbits(np.concatenate([b, np.zeros((-len(b)) % 64, dtype=bool)]).astype(np.uint8),
                                     bitorder="little").view(np.uint64)
        plans = [bdist.ProofShardPlan(ni, na, m, pack(a_bits), pack(bi_bits), pack(ba_bits), world, r)
                 for r in range(world)]
        # exponent ranges tile [0, na) and [0, m - 1)
        assert plans[0].aux_lo == 0 and plans[-1].aux_hi == na and plans[0].h_lo == 0 and plans[-1].h_hi == m - 1
        for p, q in zip(plans, plans[1:]):
            assert p.aux_hi == q.aux_lo and p.h_hi == q.h_lo and (p.aux_hi % 64 == 0 or p.aux_hi == na)
        # vector slices tile the full vectors
        n_a, n_b = ni + int(a_bits.sum()), int(bi_bits.sum()) + int(ba_bits.sum())
        for name, total in (("h", m - 1), ("l", na), ("a", n_a), ("b", n_b)):
            assert plans[0].vec[name][0] == 0 and plans[-1].vec[name][1] == total
            for p, q in zip(plans, plans[1:]):
                assert p.vec[name][1] == q.vec[name][0]
        # a_aux: global base index of the first dense position of each rank
        for p in plans:
            glob = ni + int(a_bits[:p.aux_lo].sum())
            assert p.vec["a"][0] + p.base_offset[1] == glob
            globb = int(bi_bits.sum()) + int(ba_bits[:p.aux_lo].sum())
            assert p.vec["b"][0] + p.base_offset[3] == globb == p.vec["b"][0] + p.base_offset[5]
