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


def _case(seed, n=203, nbases=None):
    rng = random.Random(seed)
    bits = [rng.random() < 0.6 for _ in range(n)]
    if nbases is None:
        nbases = 3 + sum(bits) + 2
    bases = [rng.randrange(1, P) for _ in range(nbases)]
    exps = [rng.choice([0, 1, rng.randrange(P)]) for _ in range(n)]
    return bases, exps, bits


def _words(bits):
    w = [0] * max(1, (len(bits) + 63) // 64)
    for i, b in enumerate(bits):
        if b:
            w[i // 64] |= 1 << (i % 64)
    return np.array(w, dtype=np.uint64)


def _worker(rank, world, port, seed, nbases, use_density, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    G = curves.Dummy
    bases, exps, bits = _case(seed, nbases=nbases)
    words = _words(bits) if use_density else None

    def partial_fn(lo, hi, first_base, dens_slice):
        d = ome.FullDensity()
        if dens_slice is not None:
            d = ome.DensityTracker()
            d.bv = [bool((int(dens_slice[i // 64]) >> (i % 64)) & 1) for i in range(hi - lo)]
        try:
            v = ome.multiexp(G, bases, first_base, d, exps[lo:hi], num_bits=16)
            return _lib.OK, int(v).to_bytes(4, "little")
        except ome.UnexpectedEof:
            return _lib.ERR_UNEXPECTED_EOF, bytes(4)
        except ome.UnexpectedIdentity:
            return _lib.ERR_UNEXPECTED_IDENTITY, bytes(4)

    fold = lambda parts: sum(int.from_bytes(p, "little") for p in parts) % P
    st, res = bdist.sharded_multiexp(partial_fn, fold, len(exps), words, 3 if use_density else 0)
    q.put((rank, st, res))
    dist.destroy_process_group()


def _run(seed, nbases, use_density, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, seed, nbases, use_density, q)) for r in range(2)]
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


def test_slicing_helpers():
    assert [bdist.shard_range(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    bits = [(i * 7) % 3 == 0 for i in range(200)]
    w = _words(bits)
    for lo in (0, 1, 63, 64, 65, 130, 200):
        assert bdist.dense_before(w, lo) == sum(bits[:lo])
    s = bdist.slice_density(w, 70, 150)
    assert [bool((int(s[i // 64]) >> (i % 64)) & 1) for i in range(80)] == bits[70:150]
    assert bdist.dense_before(None, 17) == 17 and bdist.slice_density(None, 0, 5) is None
