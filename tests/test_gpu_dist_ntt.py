"""GPU parity of the distributed four-step transform: the ranks of a world of 1, 2, 4, 8
are run one after the other on ONE GPU through the same C-ABI steps a multi-GPU run uses
(bmpc_fr_swap01_dev, bmpc_ntt_batch_dev, bmpc_ntt_fourstep_twiddle_dev, bmpc_fr_scale_pow_dev), with
the all-to-all emulated between them; the concatenated slices must equal the single-device
transform of the whole vector bit for bit (which tests/test_gpu_ntt.py pins to the oracle)."""
import threading

import numpy as np
import pytest
import torch

import bellman_mpc_b200 as bm
from bellman_mpc_b200 import dist as bdist
from oracle import domain as odomain
from oracle import fields

pytestmark = pytest.mark.gpu
F = fields.Fr
NAMES = {bm.FFT: "fft", bm.IFFT: "ifft", bm.COSET_FFT: "coset_fft", bm.ICOSET_FFT: "icoset_fft"}


def _rand_mont(n, seed):
    """uniform field elements as (n, 4) uint64 Montgomery limbs (any reduced limbs are a valid element)"""
    rs = np.random.RandomState(seed)
    a = rs.randint(0, 1 << 62, size=(n, 4), dtype=np.int64).astype(np.uint64)
    return a


def _single(worker, coeffs, op):
    d = bm.EvaluationDomain.from_coeffs(worker, coeffs)
    getattr(d, NAMES[op])(worker)
    out = d.into_coeffs()
    d.free()
    return out


def _emulated(worker, coeffs, log_m, world, op):
    plan = bdist.FourStepPlan(log_m, world)
    ops = bdist.GpuFrOps(worker)
    barrier = threading.Barrier(world)
    box, outs = [None] * world, [None] * world

    def run(rank):
        torch.cuda.set_device(0)

        def a2a(buf):
            torch.cuda.synchronize()
            box[rank] = buf
            barrier.wait()
            chunk = buf.shape[0] // world
            got = torch.cat([box[src][rank * chunk:(rank + 1) * chunk] for src in range(world)])
            torch.cuda.synchronize()
            barrier.wait()
            return got

        try:
            local = torch.from_numpy(coeffs[rank * plan.local:(rank + 1) * plan.local].view(np.int64).copy()).cuda()
            out = bdist.distributed_transform(local, plan, rank, op, ops, a2a)
            torch.cuda.synchronize()
            outs[rank] = out.cpu().numpy().view(np.uint64)
        except BaseException:
            barrier.abort()
            raise

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert all(o is not None for o in outs)
    return np.concatenate(outs)


@pytest.mark.parametrize("op", list(NAMES))
@pytest.mark.parametrize("log_m,world", [(2, 1), (4, 2), (7, 2), (10, 4), (13, 8), (16, 8)])
def test_four_step_equals_single_device(worker, log_m, world, op):
    coeffs = _rand_mont(1 << log_m, 300 + log_m)
    got = _emulated(worker, coeffs, log_m, world, op)
    assert np.array_equal(got, _single(worker, coeffs, op))


def test_four_step_against_oracle(worker):
    """one case straight against oracle.domain (domain.rs:81-125), not via the single-device path"""
    import random
    rng = random.Random(5)
    log_m, world = 8, 4
    vals = [rng.randrange(F.p) for _ in range(1 << log_m)]
    got = _emulated(worker, bm.fr_to_mont(vals), log_m, world, bm.COSET_FFT)
    o = odomain.EvaluationDomain(F, vals)
    o.coset_fft()
    assert bm.fr_from_mont(got) == o.coeffs


def test_batched_transform_and_transpose_steps(worker):
    """bmpc_ntt_batch_dev == bmpc_ntt_dev on every chunk; bmpc_fr_swap01_dev == numpy transpose"""
    ops = bdist.GpuFrOps(worker)
    log_n, batch = 9, 37
    a = _rand_mont(batch << log_n, 11)
    t = torch.from_numpy(a.view(np.int64).copy()).cuda()
    ops.ntt_batch(t, log_n, batch, False)
    torch.cuda.synchronize()
    got = t.cpu().numpy().view(np.uint64)
    for b in (0, 17, 36):
        assert np.array_equal(got[b << log_n:(b + 1) << log_n], _single(worker, a[b << log_n:(b + 1) << log_n], bm.FFT))
    for d0, d1, d2 in [(5, 7, 1), (33, 65, 1), (3, 4, 16), (16, 2, 3), (1, 9, 1)]:
        x = _rand_mont(d0 * d1 * d2, 12)
        t = torch.from_numpy(x.view(np.int64).copy()).cuda()
        y = ops.swap01(t, d0, d1, d2)
        torch.cuda.synchronize()
        exp = x.reshape(d0, d1, d2, 4).transpose(1, 0, 2, 3).reshape(-1, 4)
        assert np.array_equal(y.cpu().numpy().view(np.uint64), exp)
