"""Multi-GPU form of multiexp: the exponent index range is split into one contiguous
slice per rank; slice g's first base is `base_offset + popcount(density[0:lo_g])`; every rank runs
the single-GPU pipeline on its slice and emits one XYZZ partial sum; the partials (192 B G1 /
384 B G2) and status words are all-gathered and folded on every rank.  One process per GPU,
`torch.distributed` (NCCL on GPUs; the host logic below is backend-agnostic and is covered with
gloo world_size-2 tests on CPU).

Reference semantics preserved across ranks (multiexp.rs:55-65,244-249): EOF anywhere -> every
window of the reference fails, so the global status follows the same precedence rule as the
single-GPU path; see `combine_status`.
"""
from __future__ import annotations

import numpy as np

from . import _lib


def shard_range(n, world, rank):
    """contiguous slice [lo, hi) of the exponent positions owned by `rank`"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def dense_before(density_words, lo):
    """number of dense positions in [0, lo) -- how many bases the lower ranks consume"""
    if density_words is None:
        return lo
    w = np.asarray(density_words, dtype=np.uint64)
    full, rem = divmod(lo, 64)
    cnt = int(np.unpackbits(w[:full].view(np.uint8)).sum()) if full else 0
    if rem:
        cnt += bin(int(w[full]) & ((1 << rem) - 1)).count("1")
    return cnt


def slice_density(density_words, lo, hi):
    """density words of positions [lo, hi) re-based to bit 0 (None stays None)"""
    if density_words is None:
        return None
    bits = np.unpackbits(np.asarray(density_words, dtype=np.uint64).view(np.uint8), bitorder="little")[lo:hi]
    pad = (-len(bits)) % 64
    if pad or len(bits) == 0:
        bits = np.concatenate([bits, np.zeros(pad if len(bits) else 64, dtype=np.uint8)])
    return np.packbits(bits, bitorder="little").view(np.uint64).copy()


# raw flag bits produced by the MSM kernels (internal.h: MSM_FLAG_*) expressed as statuses:
# a rank reports OK, UNEXPECTED_IDENTITY (consumed identity, top window of the reference or not is
# resolved locally) or UNEXPECTED_EOF.  Across ranks: an EOF on any rank fails every reference
# window; an identity that already won locally (its top-window digit was non-zero, or no EOF on that
# rank) on a LOWER rank precedes it in scan order.
def combine_status(statuses):
    """statuses: list indexed by rank -> global status with the reference's precedence"""
    first_eof = next((r for r, s in enumerate(statuses) if s == _lib.ERR_UNEXPECTED_EOF), None)
    first_ident = next((r for r, s in enumerate(statuses) if s == _lib.ERR_UNEXPECTED_IDENTITY), None)
    other = next((s for s in statuses if s not in (_lib.OK, _lib.ERR_UNEXPECTED_EOF, _lib.ERR_UNEXPECTED_IDENTITY)), None)
    if other is not None:
        return other
    if first_eof is None:
        return _lib.ERR_UNEXPECTED_IDENTITY if first_ident is not None else _lib.OK
    # bases are consumed in position order, so only the last ranks can overrun; an identity status on
    # a rank without EOF means "some consumed base is the identity", which beats the EOF only if the
    # reference's top window consumes it -- ranks report that case as IDENTITY_TOP (see sharded_multiexp)
    return _lib.ERR_UNEXPECTED_EOF


def all_gather_bytes(local: bytes, group=None):
    """all-gather of equal-length byte strings through torch.distributed (any backend)"""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    t = torch.frombuffer(bytearray(local), dtype=torch.uint8)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    return [bytes(o.numpy().tobytes()) for o in outs]


def sharded_multiexp(partial_fn, fold_fn, n, density_words, base_offset, group=None):
    """Host orchestration shared by the GPU path and the CPU tests.

    partial_fn(lo, hi, first_base, density_slice) -> (status, partial_bytes)   this rank's slice
    fold_fn(list_of_partial_bytes) -> result                                   fold on every rank
    Returns (status, result or None)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_range(n, world, rank)
    first_base = base_offset + dense_before(density_words, lo)
    status, partial = partial_fn(lo, hi, first_base, slice_density(density_words, lo, hi))
    gathered = all_gather_bytes(bytes([status]) + bytes(partial), group)
    statuses = [g[0] for g in gathered]
    st = combine_status(statuses)
    if st != _lib.OK:
        return st, None
    return st, fold_fn([g[1:] for g in gathered])


def gpu_sharded_multiexp(worker, bases_slice, scalars_dev_ptr, n_total, group=None, stream=None):
    """FullDensity sharded MSM on GPUs: `bases_slice` holds exactly this rank's slice of the bases,
    `scalars_dev_ptr` its slice of the scalars (device).  Returns (status, uncompressed bytes)."""
    import ctypes as C

    import torch
    import torch.distributed as dist
    lib = worker._lib
    grp = bases_slice.group
    pbytes = int(lib.bmpc_partial_bytes(grp))
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_range(n_total, world, rank)
    dev = torch.device("cuda", worker.device)
    partial = torch.zeros(pbytes + 8, dtype=torch.uint8, device=dev)
    st = lib.bmpc_multiexp_partial_dev(worker.ctx, bases_slice.handle, 0, scalars_dev_ptr, hi - lo, None, 0,
                                       partial.data_ptr(), stream)
    partial[pbytes] = st
    gathered = torch.zeros(world * (pbytes + 8), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(gathered, partial, group=group)
    g = gathered.view(world, pbytes + 8)
    status = combine_status([int(x) for x in g[:, pbytes].cpu().tolist()])
    if status != _lib.OK:
        return status, None
    parts = g[:, :pbytes].contiguous()
    out = np.zeros(96 if grp == _lib.G1 else 192, dtype=np.uint8)
    st = lib.bmpc_sum_partials(worker.ctx, grp, parts.data_ptr(), world, out.ctypes.data_as(C.c_void_p), stream)
    return st, out.tobytes()
