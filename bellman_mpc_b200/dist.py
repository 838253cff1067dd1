"""Multi-GPU form of multiexp: the exponent index range is split into one contiguous
slice per rank; slice g's first base is `base_offset + popcount(density[0:lo_g])`; every rank runs
the single-GPU pipeline on its slice and emits one XYZZ partial sum; the partials (192 B G1 /
384 B G2) and status words are all-gathered and folded on every rank.  One process per GPU,
`torch.distributed` (NCCL on GPUs; the host logic below is backend-agnostic and is covered with
gloo world_size-2 tests on CPU).

Reference semantics preserved across ranks (multiexp.rs:55-65,244-249): EOF anywhere -> every
window of the reference fails, so the global status follows the same precedence rule as the
single-GPU path; see `combine_flags`.
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


# Raw flag bits produced by the MSM kernels (include/bellman_b200.h: BMPC_MSM_FLAG_*).  A shard
# cannot decide the multiexp's status alone: the reference reports the first error in scan order
# of its HIGHEST failing window (multiexp.rs:244-249).  An overrun (EOF) fails every window; an
# identity base only the windows whose digit consumes it, and every position that still finds a
# base precedes the overrun.  So the ranks exchange the raw words, OR them, and the rule is the
# single-GPU one (csrc/api.cu: flags_to_status): EOF and IDENT_TOP -> UnexpectedIdentity, EOF alone
# -> UnexpectedEof, no EOF and IDENT_ANY -> UnexpectedIdentity.  IDENT_TOP is computed against the
# reference's window for the WHOLE exponent vector (n_total is passed into every shard call).
FLAG_EOF, FLAG_IDENT_ANY, FLAG_IDENT_TOP = 1, 2, 4


def flags_status(flags):
    """status of a whole multiexp from the OR of all shards' flag words (== bmpc_msm_flags_status)"""
    if flags & FLAG_EOF:
        return _lib.ERR_UNEXPECTED_IDENTITY if flags & FLAG_IDENT_TOP else _lib.ERR_UNEXPECTED_EOF
    return _lib.ERR_UNEXPECTED_IDENTITY if flags & FLAG_IDENT_ANY else _lib.OK


def combine_flags(flag_words):
    """flag_words: one raw word per rank -> global status with the reference's precedence"""
    acc = 0
    for f in flag_words:
        acc |= int(f)
    return flags_status(acc)


def all_gather_bytes(local: bytes, group=None):
    """all-gather of equal-length byte strings through torch.distributed (gloo: host tensors;
    nccl: staged through the current CUDA device)"""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    t = torch.frombuffer(bytearray(local), dtype=torch.uint8)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    return [bytes(o.cpu().numpy().tobytes()) for o in outs]


def sharded_multiexp(partial_fn, fold_fn, n, density_words, base_offset, group=None):
    """Host orchestration shared by the GPU path and the CPU tests.

    partial_fn(lo, hi, first_base, density_slice, n_total) -> (rc, flags, partial_bytes)
        this rank's slice: rc = call status (OK unless the call itself failed), flags = raw
        BMPC_MSM_FLAG_* word of the slice (bmpc_multiexp_shard_dev)
    fold_fn(list_of_partial_bytes) -> result                                   fold on every rank
    Returns (status, result or None)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_range(n, world, rank)
    first_base = base_offset + dense_before(density_words, lo)
    rc, flags, partial = partial_fn(lo, hi, first_base, slice_density(density_words, lo, hi), n)
    gathered = all_gather_bytes(bytes([rc, flags]) + bytes(partial), group)
    failed = next((g[0] for g in gathered if g[0] != _lib.OK), None)
    if failed is not None:
        return failed, None
    st = combine_flags(g[1] for g in gathered)
    if st != _lib.OK:
        return st, None
    return st, fold_fn([g[2:] for g in gathered])


def gpu_sharded_multiexp(worker, bases_slice, scalars_dev_ptr, n_total, group=None, stream=None):
    """FullDensity sharded MSM on GPUs: `bases_slice` holds exactly this rank's slice of the bases,
    `scalars_dev_ptr` its slice of the scalars (device).  Returns (status, uncompressed bytes).
    The shard is only ENQUEUED (its record -- XYZZ partial + raw flag word -- stays on the device), the
    records are all-gathered on the same stream and folded with one host synchronisation; the status
    is the reference's for the whole vector (flag words ORed, multiexp.rs:244-249).  `stream` must be
    torch's current stream (or None with torch on the context's stream): the all-gather is ordered
    after the shard through it."""
    import ctypes as C

    import torch
    import torch.distributed as dist
    lib = worker._lib
    grp = bases_slice.group
    rbytes = int(lib.bmpc_shard_record_bytes(grp))
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_range(n_total, world, rank)
    dev = torch.device("cuda", worker.device)
    record = torch.zeros(rbytes, dtype=torch.uint8, device=dev)
    rc = lib.bmpc_multiexp_shard_enqueue_dev(worker.ctx, bases_slice.handle, 0, scalars_dev_ptr, hi - lo, None, 0,
                                             n_total, record.data_ptr(), stream)
    if rc != _lib.OK:
        # the call itself failed (arguments, CUDA): every rank must still reach the collective; the
        # failing rank's code travels in the high half of its flag word
        record.zero_()
        record[rbytes - 16 + 2] = rc          # flag word = 0x80000000 | rc << 16 (little-endian bytes)
        record[rbytes - 16 + 3] = 0x80
    gathered = torch.zeros(world * rbytes, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(gathered, record, group=group)
    out = np.zeros(96 if grp == _lib.G1 else 192, dtype=np.uint8)
    flags = C.c_uint32(0)
    st = lib.bmpc_fold_shard_records(worker.ctx, grp, gathered.data_ptr(), world, rbytes,
                                     out.ctypes.data_as(C.c_void_p), C.byref(flags), stream)
    if st != _lib.OK:
        return st, None
    if flags.value & 0x80000000:             # some rank's call failed outright
        return ((flags.value >> 16) & 0x7f) or _lib.ERR_INVALID, None
    status = flags_status(flags.value)
    if status != _lib.OK:
        return status, None
    return _lib.OK, out.tobytes()


# ------------------------------------------------------------------ create_proof over N GPUs
# every large multiexp of prover.rs:233,252-307 is sharded by exponent range; rank g
# keeps only the matching slice of each query vector.  The aux positions are cut at multiples of 64
# so that a rank's density words are a plain sub-array; the (few) input positions all go to rank 0.
JOBS = ("a_inputs", "a_aux", "b_g1_inputs", "b_g1_aux", "b_g2_inputs", "b_g2_aux", "h", "l")


def _popcount_words(words, nbits):
    return dense_before(words, nbits)


class ProofShardPlan:
    """Which slice of create_proof rank `rank` of `world` runs, and which slice [start, end) of each
    query vector (h, l, a, b_g1 == b_g2) it must hold.  Densities are fixed per circuit
    (prover.rs:100-138 sets them during synthesis from the constraint system alone), so the plan
    -- and with it the CRS distribution -- is computed once per circuit."""

    def __init__(self, num_inputs, num_aux, m, a_aux_words, b_input_words, b_aux_words, world, rank):
        self.world, self.rank = world, rank
        self.num_inputs, self.num_aux, self.m = num_inputs, num_aux, m
        blocks = (num_aux + 63) // 64
        blo, bhi = shard_range(blocks, world, rank)
        self.aux_lo, self.aux_hi = min(blo * 64, num_aux), min(bhi * 64, num_aux)
        self.in_lo, self.in_hi = (0, num_inputs) if rank == 0 else (0, 0)
        self.h_lo, self.h_hi = shard_range(m - 1, world, rank)
        a_first = num_inputs + dense_before(a_aux_words, self.aux_lo)
        a_cnt = dense_before(a_aux_words, self.aux_hi) - dense_before(a_aux_words, self.aux_lo)
        b_in_total = _popcount_words(b_input_words, num_inputs)
        b_first = b_in_total + dense_before(b_aux_words, self.aux_lo)
        b_cnt = dense_before(b_aux_words, self.aux_hi) - dense_before(b_aux_words, self.aux_lo)
        # slices of the full query vectors held by this rank
        self.vec = {
            "h": (self.h_lo, self.h_hi),
            "l": (self.aux_lo, self.aux_hi),
            "a": (0 if rank == 0 else a_first, a_first + a_cnt),
            "b": (0 if rank == 0 else b_first, b_first + b_cnt),
        }
        a0, b0 = self.vec["a"][0], self.vec["b"][0]
        # first base of each multiexp inside the rank's slice (JOBS order)
        self.base_offset = [0, a_first - a0, 0, b_first - b0, 0, b_first - b0, 0, 0]

    def shard_struct(self):
        sh = _lib.ProofShard()
        for j, v in enumerate(self.base_offset):
            sh.base_offset[j] = v
        sh.h_lo, sh.h_hi = self.h_lo, self.h_hi
        # lengths of the WHOLE exponent vectors (JOBS order): the reference's window follows them
        ni, na = self.num_inputs, self.num_aux
        for j, v in enumerate((ni, na, ni, na, ni, na, self.m - 1, na)):
            sh.n_total[j] = v
        return sh


def proof_partials(worker, params_slice, assignment, plan, skip_h=False):
    """One rank's share: (1920 partial-sum bytes, [8 raw flag words]), or (None, call status) if the
    call itself failed.  `assignment` is the FULL ProvingAssignment (host); the slices are taken here.
    skip_h: leave the H multiexp (and with it the H pipeline and the a, b, c upload) out -- its partial
    sum comes from `HSplit` then."""
    import ctypes as C
    asg = assignment
    w = worker
    s = _lib.Assignment()
    wlo = plan.aux_lo // 64
    da = np.ascontiguousarray(asg.a_aux_density.words()[wlo:])
    db = np.ascontiguousarray(asg.b_aux_density.words()[wlo:])
    dbi = np.ascontiguousarray(asg.b_input_density.words())
    aux = asg.aux_assignment[plan.aux_lo:plan.aux_hi]
    inp = asg.input_assignment[plan.in_lo:plan.in_hi]
    one = np.zeros(1, dtype=np.uint64)
    ptr = lambda a: C.c_void_p((a if a.size else one).ctypes.data)
    s.a, s.b, s.c = ptr(asg.a), ptr(asg.b), ptr(asg.c)
    s.num_constraints = asg.a.shape[0]
    s.input_assignment, s.num_inputs = ptr(inp), inp.shape[0]
    s.aux_assignment, s.num_aux = ptr(aux), aux.shape[0]
    s.a_aux_density, s.b_input_density, s.b_aux_density = ptr(da), ptr(dbi), ptr(db)
    p = params_slice._struct()
    sh = plan.shard_struct()
    if skip_h:
        sh.h_hi = sh.h_lo
    out = np.zeros(_lib.PROOF_PARTIAL_BYTES, dtype=np.uint8)
    fl = (C.c_uint32 * 8)()
    rc = w._lib.bmpc_create_proof_partials(w.ctx, C.byref(p), C.byref(s), C.byref(sh),
                                           C.c_void_p(out.ctypes.data), C.byref(fl))
    if rc != _lib.OK:
        return None, rc
    return out.tobytes(), [int(x) for x in fl]


def proof_finish(worker, params, gathered_partials, flags_by_rank, r_mont, s_mont):
    """Fold the ranks' partial sums and run the tail (prover.rs:309-349).  Returns (status, proof):
    the subversion check on delta comes first, then the multiexp statuses in the order the
    reference awaits them (prover.rs:328-343), each from the OR of the ranks' flag words."""
    import ctypes as C
    w = worker
    if (params.delta_g1[0] & 0x40) or (params.delta_g2[0] & 0x40):
        return _lib.ERR_UNEXPECTED_IDENTITY, None
    for j in range(8):
        st = combine_flags(ranks[j] for ranks in flags_by_rank)
        if st != _lib.OK:
            return st, None
    world = len(flags_by_rank)
    blob = np.frombuffer(b"".join(gathered_partials), dtype=np.uint8)
    assert blob.size == world * _lib.PROOF_PARTIAL_BYTES
    p = params._struct()
    r = np.ascontiguousarray(r_mont, dtype=np.uint64).reshape(4)
    sv = np.ascontiguousarray(s_mont, dtype=np.uint64).reshape(4)
    out = np.zeros(192, dtype=np.uint8)
    rc = w._lib.bmpc_create_proof_finish(w.ctx, C.byref(p), C.c_void_p(blob.ctypes.data), world,
                                         C.c_void_p(r.ctypes.data), C.c_void_p(sv.ctypes.data),
                                         C.c_void_p(out.ctypes.data))
    return rc, (out.tobytes() if rc == _lib.OK else None)


# The H polynomial shared by the ranks ("compute H once and scatter it").  Recomputing it
# on every rank costs each of them the whole a, b, c upload (3 x 32 m bytes) and seven transforms
# before its H multiexp can start: 17 ms of a 30 ms share at 2^22 on 8 GPUs.  Instead the ranks
# 0 .. min(world, 3) - 1 each take one of a, b, c (ifft + coset_fft of that vector only), the owners of
# b and c send their coset evaluations to rank 0 (32 m bytes each over NVLink), rank 0 finishes the
# pipeline (a b - c, divide by Z, icoset_fft, to_le_bits) and sends every rank its slice of the m - 1
# H scalars.  The seven multiexps that do not need H run meanwhile on every rank.
H_PARTIAL_OFFSET = 4 * 192          # part_g1[4] = h inside the 1920-byte blob (a_in, a_aux, b1_in, b1_aux, h, l)


def h_owner(k, world):
    """rank that transforms vector k of (a, b, c)"""
    return k % min(world, 3)


class HSplit:
    """The per-rank steps of the shared H pipeline; the exchange between them is the caller's
    (torch.distributed send / recv in `create_proof_sharded`, plain lists in the emulated tests).
    `worker`: a context of its own for these steps -- the rank's main context is busy (and locked)
    with the other seven multiexps at the same time."""

    def __init__(self, worker, plan, stream=None):
        self.w, self.plan, self.stream = worker, plan, stream
        self.log_m = plan.m.bit_length() - 1
        assert 1 << self.log_m == plan.m

    def coset_evals(self, host_vec, out_tensor):
        """out_tensor (m x 4 int64, device) <- the coset evaluations of the polynomial through the
        num_constraints host evaluations `host_vec` (pinned memory keeps the upload asynchronous)"""
        import ctypes as C
        hv = np.ascontiguousarray(host_vec, dtype=np.uint64).reshape(-1, 4)
        rc = self.w._lib.bmpc_h_coset_evals_dev(self.w.ctx, out_tensor.data_ptr(), self.log_m,
                                                C.c_void_p(hv.ctypes.data), hv.shape[0], self.stream)
        assert rc == _lib.OK, (rc, self.w._lib.bmpc_last_error(self.w.ctx))
        return out_tensor

    def combine(self, ea, eb, ec):
        """rank 0: ea <- canonical H scalars (entries 0 .. m-2) from the three coset evaluations"""
        rc = self.w._lib.bmpc_h_from_coset_evals_dev(self.w.ctx, ea.data_ptr(), eb.data_ptr(), ec.data_ptr(),
                                                     self.log_m, self.stream)
        assert rc == _lib.OK, (rc, self.w._lib.bmpc_last_error(self.w.ctx))
        return ea

    def h_partial(self, h_bases_slice, h_slice, partial_tensor):
        """this rank's share of the H multiexp over its slice of the scalars (device tensor):
        (call status, raw flag word); the XYZZ partial sum is left in partial_tensor (device)"""
        import ctypes as C
        fl = C.c_uint32(0)
        n = self.plan.h_hi - self.plan.h_lo
        rc = self.w._lib.bmpc_multiexp_shard_dev(self.w.ctx, h_bases_slice.handle, 0, h_slice.data_ptr() if n else None, n,
                                                 None, 0, self.plan.m - 1, partial_tensor.data_ptr(), C.byref(fl),
                                                 self.stream)
        return rc, int(fl.value)


def splice_h(partials, flags, h_partial_bytes, h_flags):
    """put the separately computed H partial sum and flag word into a rank's blob"""
    b = bytearray(partials)
    b[H_PARTIAL_OFFSET:H_PARTIAL_OFFSET + 192] = h_partial_bytes
    f = list(flags)
    f[6] = h_flags
    return bytes(b), f


def create_proof_sharded(assignment, params_slice, r_mont, s_mont, plan, group=None, h_worker=None):
    """create_proof (prover.rs:206-350) on all ranks of `group`: every rank calls this with its
    slice of the parameters; every rank gets the 192-byte proof.  One all-gather of 1928 bytes.
    h_worker (a second context on the rank's GPU): share the H pipeline between the ranks (`HSplit`)
    instead of recomputing it everywhere."""
    w = params_slice.worker
    if h_worker is None or plan.world < 2:
        partial, flags = proof_partials(w, params_slice, assignment, plan)
    else:
        import threading

        import torch
        import torch.distributed as dist
        rank, world, m = plan.rank, plan.world, plan.m
        dev = torch.device("cuda", w.device)
        box = {}
        th = threading.Thread(target=lambda: box.update(r=proof_partials(w, params_slice, assignment, plan, skip_h=True)))
        th.start()                                        # the seven multiexps that need no H
        side = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(side):
            hs = HSplit(h_worker, plan, side.cuda_stream)
            vecs = (assignment.a, assignment.b, assignment.c)
            ev = [None, None, None]
            for k in range(3):
                if h_owner(k, world) == rank:
                    ev[k] = hs.coset_evals(vecs[k], torch.empty((m, 4), dtype=torch.int64, device=dev))
            for k in range(3):                            # owners of b, c (and of a, if not rank 0) -> rank 0
                o = h_owner(k, world)
                if o != 0 and rank == o:
                    dist.send(ev[k], dst=0, group=group)
                if o != 0 and rank == 0:
                    ev[k] = torch.empty((m, 4), dtype=torch.int64, device=dev)
                    dist.recv(ev[k], src=o, group=group)
            n_h = plan.h_hi - plan.h_lo
            if rank == 0:
                h_all = hs.combine(ev[0], ev[1], ev[2])
                ops = []
                for g in range(1, world):
                    lo, hi = shard_range(m - 1, world, g)
                    if hi > lo:
                        ops.append(dist.P2POp(dist.isend, h_all[lo:hi], g, group=group))
                if ops:
                    for req in dist.batch_isend_irecv(ops):
                        req.wait()
                h_slice = h_all[plan.h_lo:plan.h_hi]
            else:
                h_slice = torch.empty((max(n_h, 1), 4), dtype=torch.int64, device=dev)
                if n_h:
                    for req in dist.batch_isend_irecv([dist.P2POp(dist.irecv, h_slice[:n_h], 0, group=group)]):
                        req.wait()
            hp = torch.zeros(192, dtype=torch.uint8, device=dev)
            rc_h, fl_h = hs.h_partial(params_slice.h, h_slice, hp)
            hp_bytes = bytes(hp.cpu().numpy().tobytes())
        th.join()
        partial, flags = box["r"]
        if partial is not None:
            if rc_h != _lib.OK:
                partial, flags = None, rc_h
            else:
                partial, flags = splice_h(partial, flags, hp_bytes, fl_h)
    rc = _lib.OK
    if partial is None:
        rc, partial, flags = flags, bytes(_lib.PROOF_PARTIAL_BYTES), [0] * 8
    gathered = all_gather_bytes(partial + bytes(flags) + bytes([rc]), group)
    nb = _lib.PROOF_PARTIAL_BYTES
    failed = next((g[nb + 8] for g in gathered if g[nb + 8] != _lib.OK), None)
    if failed is not None:
        return failed, None
    return proof_finish(params_slice.worker, params_slice, [g[:nb] for g in gathered],
                        [list(g[nb:nb + 8]) for g in gathered], r_mont, s_mont)


# ----------------------------------------------------------------------------------------------
# Distributed EvaluationDomain transform ("four-step NTT with an NCCL all-to-all
# transpose"): the m = 2^log_m coefficients are split over G ranks, rank g holding the contiguous
# slice [g m/G, (g+1) m/G), and fft / ifft / coset_fft / icoset_fft (src/domain.rs:81-125) produce
# the same slice of the same output vector as best_fft on one device.
#
# With m = R C, j = C j1 + j2 and k = k1 + R k2:
#   X[k1 + R k2] = sum_j2 w_C^(j2 k2) * [ w_m^(j2 k1) * sum_j1 x[C j1 + j2] w_R^(j1 k1) ]
# i.e. (T) transpose the R x C matrix so every rank holds whole columns, R-point transforms of the
# columns, twiddle by w_m^(j2 k1), (T) transpose so every rank holds whole rows k1, C-point
# transforms, (T) transpose into natural order.  Each (T) is: local block permutation, one
# all-to-all of equal chunks, local transpose.  The local steps are C-ABI calls (`GpuFrOps`); the
# exchange is `torch.distributed.all_to_all_single` (NCCL over NVLink; gloo in the CPU tests).

def is_pow2(x):
    return x >= 1 and (x & (x - 1)) == 0


class FourStepPlan:
    """geometry of the distributed transform: m = R x C with R = 2^ceil(log_m / 2)"""

    def __init__(self, log_m, world):
        if not is_pow2(world):
            raise ValueError("distributed transform needs a power-of-two number of ranks")
        self.log_m, self.world = log_m, world
        self.log_r = (log_m + 1) // 2
        self.log_c = log_m - self.log_r
        self.R, self.C = 1 << self.log_r, 1 << self.log_c
        if world > self.C or self.log_c < 1:
            raise ValueError(f"domain 2^{log_m} is too small to split over {world} ranks")
        self.m = 1 << log_m
        self.local = self.m // world          # coefficients per rank


def _dist_transpose(buf, P, Q, G, ops, all_to_all):
    """buf = this rank's P/G rows of a P x Q matrix (row-major) -> its Q/G rows of the transpose"""
    t = ops.swap01(buf, P // G, G, Q // G)        # [P/G][G][Q/G] -> [G][P/G][Q/G]: chunk h goes to rank h
    r = all_to_all(t) if G > 1 else t             # [G (source)][P/G][Q/G] = all P rows of my Q/G columns
    return ops.swap01(r, P, Q // G, 1)            # -> [Q/G][P]


def distributed_transform(local, plan, rank, op, ops, all_to_all):
    """One of FFT / IFFT / COSET_FFT / ICOSET_FFT over a domain split across `plan.world` ranks.
    local: this rank's slice in the buffer type of `ops`; returns the rank's slice of the result.
    ops: local steps (GpuFrOps, or the oracle-backed ops of the CPU tests); all_to_all(buf) -> buf
    exchanges `world` equal chunks (chunk h of the input goes to rank h)."""
    G, R, C = plan.world, plan.R, plan.C
    inverse = op in (_lib.IFFT, _lib.ICOSET_FFT)
    first = rank * plan.local
    if op == _lib.COSET_FFT:
        ops.scale_pow(local, plan.local, first, plan.log_m, 0)           # distribute_powers(g), :116-119
    b = _dist_transpose(local, R, C, G, ops, all_to_all)                 # C/G columns of length R
    ops.ntt_batch(b, plan.log_r, C // G, inverse)
    ops.twiddle(b, C // G, R, rank * (C // G), plan.log_m, inverse)
    d = _dist_transpose(b, C, R, G, ops, all_to_all)                     # R/G rows k1 of length C
    ops.ntt_batch(d, plan.log_c, R // G, inverse)
    out = _dist_transpose(d, R, C, G, ops, all_to_all)                   # X[k1 + R k2] at row k2, column k1
    if op == _lib.IFFT:
        ops.scale_pow(out, plan.local, first, plan.log_m, 2)             # minv, :88-98
    elif op == _lib.ICOSET_FFT:
        ops.scale_pow(out, plan.local, first, plan.log_m, 1)             # g^-i / m, :121-125
    return out


class GpuFrOps:
    """the local steps on this rank's GPU through the C ABI; buffers are torch CUDA tensors of shape
    (n, 4) uint64 (Montgomery limbs) -- torch only owns the memory and the stream"""

    def __init__(self, worker):
        import torch
        self.torch = torch
        self.lib, self.ctx = worker._lib, worker.ctx

    def _stream(self):
        # torch's current stream, so the steps are ordered with torch's copies and with NCCL.  The C
        # ABI reads a NULL stream as "the context's own stream", so torch's default stream (handle 0)
        # is passed as cudaStreamLegacy (0x1), the explicit name of the same stream.
        s = self.torch.cuda.current_stream().cuda_stream
        return s if s else 1

    def _ck(self, rc):
        if rc != _lib.OK:
            raise RuntimeError(f"distributed transform step failed: status {rc}: "
                               f"{(self.lib.bmpc_last_error(self.ctx) or b'').decode()}")

    def swap01(self, t, d0, d1, d2):
        out = self.torch.empty_like(t)
        self._ck(self.lib.bmpc_fr_swap01_dev(self.ctx, t.data_ptr(), out.data_ptr(), d0, d1, d2, self._stream()))
        return out

    def ntt_batch(self, t, log_n, batch, inverse):
        self._ck(self.lib.bmpc_ntt_batch_dev(self.ctx, t.data_ptr(), log_n, batch, int(inverse), self._stream()))

    def twiddle(self, t, rows, cols, row0, log_m, inverse):
        self._ck(self.lib.bmpc_ntt_fourstep_twiddle_dev(self.ctx, t.data_ptr(), rows, cols, row0, log_m,
                                                        int(inverse), self._stream()))

    def scale_pow(self, t, n, first, log_m, which):
        self._ck(self.lib.bmpc_fr_scale_pow_dev(self.ctx, t.data_ptr(), n, first, log_m, which, self._stream()))


def torch_all_to_all(group=None):
    """all_to_all callable for `distributed_transform` over torch.distributed"""
    import torch
    import torch.distributed as dist

    def run(t):
        out = torch.empty_like(t)
        dist.all_to_all_single(out, t, group=group)
        return out
    return run


class DistributedDomain:
    """EvaluationDomain (domain.rs:21-125) whose coefficients are split over the ranks of a
    torch.distributed group: same method names, each rank passes and gets back its contiguous slice."""

    def __init__(self, worker, local, log_m, rank=None, world=None, group=None):
        import torch.distributed as dist
        self.world = dist.get_world_size(group) if world is None else world
        self.rank = dist.get_rank(group) if rank is None else rank
        self.plan = FourStepPlan(log_m, self.world)
        if local.shape[0] != self.plan.local:
            raise AssertionError("local slice length does not match the domain / world size")
        self.local, self.ops = local, GpuFrOps(worker)
        self.a2a = torch_all_to_all(group)

    def _t(self, op):
        self.local = distributed_transform(self.local, self.plan, self.rank, op, self.ops, self.a2a)

    def fft(self, worker=None):
        self._t(_lib.FFT)

    def ifft(self, worker=None):
        self._t(_lib.IFFT)

    def coset_fft(self, worker=None):
        self._t(_lib.COSET_FFT)

    def icoset_fft(self, worker=None):
        self._t(_lib.ICOSET_FFT)

    def into_coeffs(self):
        return self.local
