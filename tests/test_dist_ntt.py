"""Host logic of the distributed four-step transform (bellman_mpc_b200/dist.py:
distributed_transform) on the CPU: the local steps are restated over oracle.fields.Fr
integers, the exchange is either emulated in-process (world 1, 2, 4, 8) or a real
`torch.distributed.all_to_all_single` over gloo (world 2); the rank slices must concatenate to
exactly what oracle.domain.EvaluationDomain (domain.rs:81-125) gives on the whole vector."""
import os
import random

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bellman_mpc_b200 import _lib
from bellman_mpc_b200 import dist as bdist
from oracle import domain as odom
from oracle import fields

F = fields.Fr
OPS = {_lib.FFT: "fft", _lib.IFFT: "ifft", _lib.COSET_FFT: "coset_fft", _lib.ICOSET_FFT: "icoset_fft"}


def _omega(log_n):
    w = F.root_of_unity
    for _ in range(log_n, F.S):
        w = w * w % F.p
    return w


class OracleFrOps:
    """bmpc_fr_swap01_dev / bmpc_ntt_batch_dev / bmpc_ntt_fourstep_twiddle_dev / bmpc_fr_scale_pow_dev
    restated over python integers (buffers are lists)"""

    def swap01(self, buf, d0, d1, d2):
        out = [0] * len(buf)
        for a in range(d0):
            for b in range(d1):
                src, dst = (a * d1 + b) * d2, (b * d0 + a) * d2
                out[dst:dst + d2] = buf[src:src + d2]
        return out

    def ntt_batch(self, buf, log_n, batch, inverse):
        n = 1 << log_n
        w = _omega(log_n)
        if inverse:
            w = F.inv(w)
        for t in range(batch):
            chunk = buf[t * n:(t + 1) * n]
            odom.best_fft(F, chunk, w, log_n)
            buf[t * n:(t + 1) * n] = chunk

    def twiddle(self, buf, rows, cols, row0, log_m, inverse):
        w = _omega(log_m)
        if inverse:
            w = F.inv(w)
        for r in range(rows):
            for c in range(cols):
                buf[r * cols + c] = buf[r * cols + c] * pow(w, (row0 + r) * c, F.p) % F.p

    def scale_pow(self, buf, n, first, log_m, which):
        g = F.generator
        minv = F.inv((1 << log_m) % F.p)
        for i in range(n):
            if which == 0:
                f = pow(g, first + i, F.p)
            elif which == 1:
                f = pow(F.inv(g), first + i, F.p) * minv % F.p
            else:
                f = minv
            buf[i] = buf[i] * f % F.p


def _expected(coeffs, op):
    d = odom.EvaluationDomain(F, coeffs)
    getattr(d, OPS[op])()
    return d.into_coeffs()


def _emulated(coeffs, log_m, world, op):
    """all ranks in one process: the exchange waits until every rank has produced its chunks"""
    plan = bdist.FourStepPlan(log_m, world)
    ops = OracleFrOps()
    # run the ranks as generators that yield at every all-to-all
    import threading
    barrier = threading.Barrier(world)
    box = [None] * world
    outs = [None] * world

    def run(rank):
        def a2a(buf):
            box[rank] = buf
            barrier.wait()
            chunk = len(buf) // world
            got = []
            for src in range(world):
                got.extend(box[src][rank * chunk:(rank + 1) * chunk])
            barrier.wait()
            return got
        local = list(coeffs[rank * plan.local:(rank + 1) * plan.local])
        outs[rank] = bdist.distributed_transform(local, plan, rank, op, ops, a2a)

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    return [v for o in outs for v in o]


@pytest.mark.parametrize("op", list(OPS))
@pytest.mark.parametrize("log_m,world", [(2, 1), (3, 2), (5, 2), (6, 4), (7, 8), (8, 4)])
def test_four_step_matches_the_reference_transform(log_m, world, op):
    rng = random.Random(100 * log_m + world)
    coeffs = [rng.randrange(F.p) for _ in range(1 << log_m)]
    assert _emulated(coeffs, log_m, world, op) == _expected(coeffs, op)


def test_plan_rejects_bad_geometry():
    with pytest.raises(ValueError):
        bdist.FourStepPlan(10, 3)
    with pytest.raises(ValueError):
        bdist.FourStepPlan(3, 4)          # C = 2 columns cannot be split over 4 ranks


def _limbs(vals):
    return np.array([[(v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)] for v in vals], dtype=np.uint64)


def _gloo_worker(rank, world, port, log_m, op, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = random.Random(7)
    coeffs = [rng.randrange(F.p) for _ in range(1 << log_m)]
    plan = bdist.FourStepPlan(log_m, world)

    def a2a(buf):                         # lists of integers <-> (n, 4) uint64 tensors on the wire
        t = torch.from_numpy(_limbs(buf).view(np.int64))
        out = torch.empty_like(t)
        dist.all_to_all_single(out, t)
        arr = out.numpy().view(np.uint64)
        return [sum(int(arr[i, j]) << (64 * j) for j in range(4)) for i in range(arr.shape[0])]

    local = list(coeffs[rank * plan.local:(rank + 1) * plan.local])
    got = bdist.distributed_transform(local, plan, rank, op, OracleFrOps(), a2a)
    q.put((rank, got))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("op", [_lib.FFT, _lib.ICOSET_FFT])
def test_four_step_over_gloo_world2(op):
    log_m, world = 6, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + random.randrange(300)
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, log_m, op, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = random.Random(7)
    coeffs = [rng.randrange(F.p) for _ in range(1 << log_m)]
    assert res[0] + res[1] == _expected(coeffs, op)
