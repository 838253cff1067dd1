"""Synthetic Groth16 proving workload (BASELINE config #4 / SURVEY 8d-4) for bench.py and the GPU
tests: the post-synthesis state of a 2^k-constraint R1CS with random density -- evaluation vectors
a, b, c (c = a o b so the quotient is exact), 16 inputs, 2^k - 16 aux variables, A-aux density 0.75,
B densities 0.5 -- and a CRS whose points have KNOWN discrete logs, so the whole create_proof
pipeline can be checked bit-for-bit at any size with O(n) field arithmetic on the CPU
(SURVEY 8c "known-trapdoor checks") instead of a CPU MSM.

Timed region = src/groth16/prover.rs:206-350 (everything after synthesis); synthesis itself stays
on the CPU in the reference and is not part of the hot path.
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np

import bellman_mpc_b200 as bm

Q = bm.FR_MODULUS
G1_GEN = bytes.fromhex(
    "17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb"
    "08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1")
G2_GEN = bytes.fromhex(
    "13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e"
    "024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8"
    "0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be"
    "0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801")
# the fork's fixed protocol scalars: generator.rs:34-38, prover.rs:169-170
ALPHA, BETA, DELTA, R_, S_ = 6, 24, 24, 27134, 17146


def rand_limbs(n, seed):
    rs = np.random.RandomState(seed)
    a = rs.randint(0, 1 << 63, size=(n, 4), dtype=np.int64).astype(np.uint64)
    a[:, 3] >>= np.uint64(1)
    return a


def pinned(arr):
    """copy into page-locked host memory (the H2D copies inside the timed region start there)"""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(arr).view(np.int64)).pin_memory()
    return t.numpy().view(np.uint64), t


class Workload:
    def __init__(self, w, log_m, seed=4, profile="uniform", precompute=False, world=1, rank=0):
        """world > 1: this process is rank `rank` of a sharded prover and registers only
        its slices of the query vectors (bellman_mpc_b200.dist.ProofShardPlan)."""
        self.w, self.log_m = w, log_m
        self.world, self.rank = world, rank
        m = 1 << log_m
        ni = 16 if m > 32 else 2
        na = m - ni
        self.m, self.ni, self.na = m, ni, na
        rs = np.random.RandomState(seed)
        # evaluations (Montgomery limbs of uniform values); c = a o b computed on the GPU
        a, b = rand_limbs(m, seed + 10), rand_limbs(m, seed + 11)
        da, db = bm.EvaluationDomain.from_coeffs(w, a), bm.EvaluationDomain.from_coeffs(w, b)
        da.mul_assign(w, db)
        c = da.into_coeffs()
        da.free(); db.free()
        if profile == "boolean":      # boolean-heavy witness: 0 / 1 / uniform (Montgomery forms)
            vals = rand_limbs(ni + na, seed + 12)
            kind = rs.randint(0, 10, size=ni + na)
            one = bm.fr_to_mont([1])[0]
            vals[kind < 4] = 0
            vals[(kind >= 4) & (kind < 8)] = one
        else:
            vals = rand_limbs(ni + na, seed + 12)
        vals[0] = bm.fr_to_mont([1])[0]                      # ONE
        self._keep = []
        (self.a, t1), (self.b, t2), (self.c, t3) = pinned(a), pinned(b), pinned(c)
        (self.inputs, t4), (self.aux, t5) = pinned(vals[:ni]), pinned(vals[ni:])
        self._keep += [t1, t2, t3, t4, t5]
        self.a_aux_bits = rs.random_sample(na) < 0.75
        self.b_in_bits = rs.random_sample(ni) < 0.5
        self.b_in_bits[0] = True
        self.b_aux_bits = rs.random_sample(na) < 0.5
        self.dens = [bm.DensityTracker.from_bits(x) for x in (self.a_aux_bits, self.b_in_bits, self.b_aux_bits)]
        # CRS with known discrete logs
        n_a = ni + int(self.a_aux_bits.sum())
        n_b = int(self.b_in_bits.sum()) + int(self.b_aux_bits.sum())
        self.k_h, self.k_l = rand_limbs(m - 1, seed + 20), rand_limbs(na, seed + 21)
        self.k_a, self.k_b = rand_limbs(n_a, seed + 22), rand_limbs(n_b, seed + 23)
        from bellman_mpc_b200 import dist as bdist
        self.plan = bdist.ProofShardPlan(ni, na, m, self.dens[0].words(), self.dens[1].words(), self.dens[2].words(),
                                         world, rank)
        sl = lambda k, name: np.ascontiguousarray(k[self.plan.vec[name][0]:self.plan.vec[name][1]])
        one = lambda k: k if len(k) else rand_limbs(1, 1)      # a rank may own an empty slice
        fb = lambda wk, grp, gen, k: bm.Bases.fixed_base_mul(wk, grp, gen, one(k))
        t0 = time.perf_counter()
        self.h, self.l = fb(w, bm.G1, G1_GEN, sl(self.k_h, "h")), fb(w, bm.G1, G1_GEN, sl(self.k_l, "l"))
        self.qa, self.qb1 = fb(w, bm.G1, G1_GEN, sl(self.k_a, "a")), fb(w, bm.G1, G1_GEN, sl(self.k_b, "b"))
        self.qb2 = fb(w, bm.G2, G2_GEN, sl(self.k_b, "b"))
        fb = bm.Bases.fixed_base_mul
        vk1 = fb(w, bm.G1, G1_GEN, bm.ints_to_limbs([ALPHA, BETA, DELTA])).read()
        vk2 = fb(w, bm.G2, G2_GEN, bm.ints_to_limbs([BETA, DELTA])).read()
        if precompute:
            for q in (self.h, self.l, self.qa, self.qb1, self.qb2):
                q.precompute()
        self.crs_setup_s = time.perf_counter() - t0
        self.params = bm.Parameters(w, self.h, self.l, self.qa, self.qb1, self.qb2, vk1[:96], vk1[96:192],
                                    vk2[:192], vk1[192:288], vk2[192:384])
        self.params.gamma_g2 = vk2[:192]                       # gamma = beta in this synthetic vk
        self.params.ic = fb(w, bm.G1, G1_GEN, rand_limbs(ni, seed + 24))
        self.assignment = bm.ProvingAssignment(self.a, self.b, self.c, self.inputs, self.aux, *self.dens)
        self.r = bm.fr_to_mont([R_])[0]
        self.s = bm.fr_to_mont([S_])[0]

    def prove(self):
        if self.world > 1:
            from bellman_mpc_b200 import dist as bdist
            st, proof = bdist.create_proof_sharded(self.assignment, self.params, self.r, self.s, self.plan,
                                                   h_worker=getattr(self, "h_worker", None))
            assert st == 0, st
            return proof
        return bm.create_proof(self.assignment, self.params, self.r, self.s)

    def h2d_bytes(self):
        p = self.plan
        vecs = 3                      # a, b, c uploaded by this rank
        if self.world > 1 and getattr(self, "h_worker", None) is not None:
            from bellman_mpc_b200 import dist as bdist
            vecs = sum(1 for k in range(3) if bdist.h_owner(k, self.world) == self.rank)
        return (32 * (vecs * self.m + (p.in_hi - p.in_lo) + (p.aux_hi - p.aux_lo)) + 3 * ((p.aux_hi - p.aux_lo + 63) // 64) * 8
                + self.world * 1928)

    def free(self):
        for b in (self.h, self.l, self.qa, self.qb1, self.qb2):
            b.free()

    # -------------------------------------------------------------- checker (uses oracle/)
    def expected_proof(self):
        """Known-dlog expectation of the proof bytes: O(n) field arithmetic + 3 scalar mults."""
        from oracle import cref, curves
        w = self.w
        to_int = lambda limbs: bm.limbs_to_ints(np.asarray(limbs).reshape(1, 4))[0]
        canon = lambda mont: fr_mont_to_canonical(w, mont)
        inputs, aux = canon(self.inputs), canon(self.aux)
        h = bm.h_coefficients(w, self.a, self.b, self.c)
        dot = lambda k, s: to_int(cref.fr_dot(k, s)) if len(k) else 0
        ni = self.ni
        a_sum = (dot(self.k_a[:ni], inputs) + dot(self.k_a[ni:], aux[self.a_aux_bits])) % Q
        nbi = int(self.b_in_bits.sum())
        b_sum = (dot(self.k_b[:nbi], inputs[self.b_in_bits]) + dot(self.k_b[nbi:], aux[self.b_aux_bits])) % Q
        h_s, l_s = dot(self.k_h, h), dot(self.k_l, aux)
        a_s = (ALPHA + a_sum + R_ * DELTA) % Q
        b_s = (BETA + b_sum + S_ * DELTA) % Q
        c_s = (R_ * S_ * DELTA + S_ * ALPHA + R_ * BETA + S_ * a_sum + R_ * b_sum + h_s + l_s) % Q
        G1, G2 = curves.G1, curves.G2
        return (G1.to_compressed(G1.mul(G1.gen, a_s)) + G2.to_compressed(G2.mul(G2.gen, b_s))
                + G1.to_compressed(G1.mul(G1.gen, c_s)))

    def cpu_reference_proof(self, threads):
        """The C restatement of prover.rs:206-350 on the host cores (timed CPU baseline)."""
        from oracle import cref
        cp = cref.CParams()
        hs = [cref.CBases.from_uncompressed(b.group, b.read()) for b in (self.h, self.l, self.qa, self.qb1, self.qb2)]
        cp.h, cp.l, cp.a, cp.b_g1, cp.b_g2 = (x.handle for x in hs)
        p = self.params
        for name in ("alpha_g1", "beta_g1", "beta_g2", "delta_g1", "delta_g2"):
            C.memmove(getattr(cp, name), getattr(p, name), len(getattr(p, name)))
        t0 = time.perf_counter()
        st, proof = cref.create_proof(cp, self.a, self.b, self.c, self.inputs, self.aux, self.dens[0].words(),
                                      self.dens[1].words(), self.dens[2].words(), self.r, self.s, threads=threads)
        dt = time.perf_counter() - t0
        for x in hs:
            x.free()
        return st, proof, dt


def fr_mont_to_canonical(w, mont):
    """PrimeFieldBits::to_le_bits on the GPU for a host array (helper for the checker)"""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(mont).view(np.int64)).cuda()
    st = w._lib.bmpc_fr_to_canonical_dev(w.ctx, t.data_ptr(), t.shape[0], torch.cuda.current_stream().cuda_stream)
    assert st == 0
    torch.cuda.synchronize()
    return t.cpu().numpy().view(np.uint64)


def prove_bench_sharded(w, log_m, steps, world, rank, precompute=True):
    """Groth16 prove on `world` GPUs (one process each): CRS slices resident per rank, the H
    polynomial recomputed on every rank, one 1928-byte all-gather, tail on every rank.  Timed with
    CUDA events around the (host-synchronous) call after a barrier, max over ranks."""
    import torch
    import torch.distributed as dist
    wl = Workload(w, log_m, precompute=precompute, world=world, rank=rank)
    # a second context on this GPU for the shared H pipeline (bellman_mpc_b200.dist.HSplit); BMPC_H_SPLIT=0:
    # every rank recomputes H (the round-1 path)
    import os
    h_split = os.environ.get("BMPC_H_SPLIT", "1") != "0"
    if h_split:
        wl.h_worker = bm.Worker(w.device)
    wl.prove()
    times = []
    proof = None
    for _ in range(steps):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        proof = wl.prove()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
    res = {
        "metric": "groth16_prove_seconds", "constraints": 1 << log_m, "value": min(times), "unit": "s",
        "mean_s": sum(times) / len(times), "all_s": [round(t, 4) for t in times], "steps": steps, "n_gpus": world,
        "higher_is_better": False,
        "timed_region": ("prover.rs:206-350 sharded: per rank its slice of the assignments -> H2D -> 7 partial MSMs, while "
                         "ranks 0-2 upload one of a, b, c each (pinned host) and transform it, rank 0 combines them into H and "
                         "sends every rank its slice -> partial H MSM -> all-gather of 1928 B -> fold + tail -> 192-byte "
                         "proof on every rank") if h_split else
                        ("prover.rs:206-350 sharded: per rank pinned host a,b,c + its slice of the assignments -> H2D -> "
                         "7 NTT + 8 partial MSMs -> all-gather of 1928 B -> fold + tail -> 192-byte proof on every rank"),
        "h_pipeline": "shared: one vector per rank 0-2, combined on rank 0, slices over NVLink" if h_split else "recomputed on every rank",
        "h2d_bytes_per_step": wl.h2d_bytes(), "d2h_bytes_per_step": 192 + 1920,
        "parallelism": f"multiexp exponent ranges split x{world} (aux positions in blocks of 64)",
        "crs_setup_s": round(wl.crs_setup_s, 2),
    }
    if rank == 0:
        res["matches_known_dlog_expectation"] = bool(proof == wl.expected_proof())
    wl.free()
    return res


def prove_bench(w, log_m, steps, no_cpu_baseline=False, cpu_sample_log=None, precompute=True):
    """cpu_sample_log None: the CPU restatement proves the SAME 2^log_m workload once (the stated
    configuration; ~40 s on 16 cores at 2^22)."""
    import torch
    wl = Workload(w, log_m, precompute=precompute)
    wl.prove()                                     # warm-up (tables, arena)
    torch.cuda.synchronize()
    launches0 = w.launch_count()
    times = []
    proof = None
    for _ in range(steps):
        t0 = time.perf_counter()
        proof = wl.prove()
        times.append(time.perf_counter() - t0)
    launches = (w.launch_count() - launches0) // steps
    res = {
        "metric": "groth16_prove_seconds", "constraints": 1 << log_m, "value": min(times), "unit": "s",
        "mean_s": sum(times) / len(times), "all_s": [round(t, 4) for t in times], "steps": steps,
        "higher_is_better": False,
        "timed_region": "prover.rs:206-350 through bmpc_create_proof: pinned host a,b,c + assignments + "
                        "densities -> H2D -> 7 NTT + 8 MSM + tail -> 192-byte proof D2H",
        "h2d_bytes_per_step": wl.h2d_bytes(), "d2h_bytes_per_step": 192,
        "workload": f"synthetic R1CS 2^{log_m} constraints, 16 inputs, A-aux density 0.75, B density 0.5, uniform witness",
        "crs_setup_s": round(wl.crs_setup_s, 2), "gpu_launches_per_proof": launches,
    }
    if not no_cpu_baseline:
        from oracle import cref
        t0 = time.perf_counter()
        res["matches_known_dlog_expectation"] = bool(proof == wl.expected_proof())
        res["check_s"] = round(time.perf_counter() - t0, 2)
        threads = cref.hardware_threads()
        if cpu_sample_log is None or cpu_sample_log >= log_m:
            st, cp, dt = wl.cpu_reference_proof(threads)
            res["cpu_baseline"] = {"value": dt, "unit": "s", "cores": threads, "kind": "port",
                                   "sample": f"the same 2^{log_m}-constraint workload, proved once (prover.rs:206-350 "
                                             "restated in oracle/c: 7 transforms + 8 multiexps, one task per window)",
                                   "matches_gpu_bytes": bool(st == 0 and cp == proof),
                                   "gpu_seconds_same_sample": min(times),
                                   "speedup": dt / min(times)}
        else:
            small = Workload(w, cpu_sample_log, precompute=precompute)
            gp = small.prove()
            st, cp, dt = small.cpu_reference_proof(threads)
            res["cpu_baseline"] = {"value": dt, "unit": "s", "cores": threads, "kind": "port",
                                   "sample": f"same generator at 2^{cpu_sample_log} constraints (full prove)",
                                   "matches_gpu_bytes": bool(st == 0 and cp == gp),
                                   "gpu_seconds_same_sample": None}
            t0 = time.perf_counter()
            small.prove()
            res["cpu_baseline"]["gpu_seconds_same_sample"] = time.perf_counter() - t0
            small.free()
    # Parameters::read ingestion on a 2^18-constraint CRS: write, then read back
    try:
        pw = Workload(w, 18)
        blob = pw.params.write()
        io = {"constraints": 1 << 18, "blob_bytes": len(blob)}
        for checked in (False, True):
            t0 = time.perf_counter()
            back = bm.Parameters.read(w, blob, checked)
            dt = time.perf_counter() - t0
            ok = back.write() == blob
            io["checked" if checked else "unchecked"] = {"seconds": dt, "gb_per_s": len(blob) / dt / 1e9,
                                                         "roundtrip_identical": bool(ok)}
            back.free()
        res["params_read"] = io
        pw.free()
    except Exception as e:
        res["params_read"] = {"error": repr(e)}
    wl.free()
    return res
