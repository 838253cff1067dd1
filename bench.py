#!/usr/bin/env python
"""bench.py -- headline benchmark of the Groth16 hot path on B200.

Metric (BASELINE.json): G1 MSM Mpts/s @ 2^24 at 1/2/4/8 B200 (the scaling metric; `value`),
with Groth16 prove seconds @ 2^22 constraints reported in the same line under "prove" (N = 1).

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, sm_100a)
  python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference

One "step" = one multiexp over the whole 2^24-point workload (strong scaling: the bases are
split across the N ranks, each rank reduces its slice to a partial sum, the 192-byte partials
are all-gathered over NCCL and folded on every rank).  Inputs are resident in HBM for `value`;
`e2e` times the host-buffer C-ABI call (pinned host scalars -> H2D inside the timed region,
result bytes D2H).  Bases are the CRS: registered once, resident, like `Parameters` in a prover.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "g1_msm_mpts_per_s_2p24"
UNIT = "Mpts/s"
G1_MSM_BYTES_PER_POINT = 128        # 96 B affine base + 32 B scalar
G1_MSM_MAC32_PER_POINT = 48_000     # 16 windows x 10 Fp mul x 300 MAC32
G2_MSM_MAC32_PER_POINT = 144_000


def rand_limbs(n, seed):
    """n x 4 u64, uniform below 2^254 (< q): canonical scalars"""
    rs = np.random.RandomState(seed)
    a = rs.randint(0, 1 << 63, size=(n, 4), dtype=np.int64).astype(np.uint64)
    a[:, 3] >>= np.uint64(1)
    return a


def measured_peaks():
    peaks = {"hbm_gbs": 6650.0, "source": "fallback"}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            peaks["hbm_gbs"] = float(json.load(open(p))["hbm_gbs"])
            peaks["source"] = "measured"
        except Exception:
            pass
    kp = os.path.join(ROOT, "profiles", "r01_ncu_kernel_summaries.json")
    if os.path.exists(kp):
        try:
            k = json.load(open(kp))["msm_accumulate_g1"]
            # ncu --set full capture of this bench command (2^24, 1 GPU): GB read + MB written per launch
            peaks["acc_traffic_bytes"] = float(k["dram__bytes_read.sum"]) * 1e9 + float(k["dram__bytes_write.sum"]) * 1e6
        except Exception:
            pass
        try:
            peaks["aff_traffic_bytes"] = float(json.load(open(kp))["msm_accumulate_affine_g1"]["dram_bytes_total"])
        except Exception:
            pass
    kp2 = os.path.join(ROOT, "profiles", "r02_ncu_kernel_summaries.json")     # the kernel as it is now
    if os.path.exists(kp2):
        try:
            k2 = json.load(open(kp2))["msm_accumulate_affine_g1"]
            peaks["aff_traffic_bytes"] = float(k2["dram_bytes_total"])
            peaks["fmaheavy_pct"] = float(k2["sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"])
        except Exception:
            pass
    ip = os.path.join(ROOT, "profiles", "r01_imad_peak.json")
    if os.path.exists(ip):
        try:
            peaks["mac32_per_s"] = float(json.load(open(ip))["cc_pair_mac32_per_s"])
        except Exception:
            pass
    return peaks


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------ reference arm
def dlog_progression(seed):
    """(first, step) of the reference arm's bases P_i = (first + i step) G: discrete logs spread over
    the whole scalar field, still known in closed form for the result check"""
    import random
    from oracle import fields
    rng = random.Random(seed)
    return rng.randrange(1, fields.Fr.p), rng.randrange(1, fields.Fr.p)


def limbs_to_int(l):
    return sum(int(l[j]) << (64 * j) for j in range(4))


def int_to_limbs(v):
    return np.array([(v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)], dtype=np.uint64)


def run_reference(args):
    """The reference's CPU multiexp (multiexp.rs:159-281) as restated in oracle/c (64-bit-limb
    Montgomery, one task per window as multiexp.rs:238-242, pthreads) on all host cores, at the
    STATED configuration: uniform 254-bit scalars, bases with field-wide discrete logs, and the window
    the reference picks for 2^log_n exponents (c = 17, 15 windows at 2^24).  Each step is one multiexp
    over a 2^ref_sample_log-point sample of that vector (default 2^23: the per-window bucket sums the
    sample pays once per step cost 3 % of its additions, so the per-point cost is the full vector's
    to within that), result checked against (sum k_i s_i) G."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cref, fields
    lib = cref.load()
    threads = cref.hardware_threads()
    n_full = 1 << args.log_n
    sample_log = min(args.ref_sample_log, args.log_n)
    n = 1 << sample_log
    first, step = dlog_progression(2)
    t_setup = time.perf_counter()
    bases = cref.bases_g1_progression(n, first, step)
    t_setup = time.perf_counter() - t_setup
    scalars = rand_limbs(n, 1)
    c_ref = int(lib.orc_window_size(n_full))
    for _ in range(args.warmup):
        cref.multiexp_window(bases, 0, scalars, n_full, threads=threads)
    t0 = time.perf_counter()
    out = None
    for _ in range(args.steps):
        st, out = cref.multiexp_window(bases, 0, scalars, n_full, threads=threads)
        assert st == 0
    dt = (time.perf_counter() - t0) / args.steps
    ones = np.zeros((n, 4), dtype=np.uint64)
    ones[:, 0] = 1
    idx = np.zeros((n, 4), dtype=np.uint64)
    idx[:, 0] = np.arange(n, dtype=np.uint64)
    q = fields.Fr.p
    k = (first * limbs_to_int(cref.fr_dot(ones, scalars)) + step * limbs_to_int(cref.fr_dot(idx, scalars))) % q
    ok = out == cref.g1_generator_mul(int_to_limbs(k))
    value = n / dt / 1e6
    windows = (255 + c_ref - 1) // c_ref
    sample = (f"G1 multiexp over a 2^{sample_log}-point sample of the 2^{args.log_n} vector per step, reference window "
              f"c={c_ref} ({windows} windows = {windows} tasks, multiexp.rs:238-242,267-271) as for the whole vector")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u64-limb Montgomery (CPU)", "data": "synthetic",
        "config": {"workload": f"G1 multiexp 2^{args.log_n} points, uniform 254-bit scalars, FullDensity "
                               f"(BASELINE configs[1] shape at the size the metric is quoted on)",
                   "bases": "(a + i b) G with random 255-bit a, b (field-wide known discrete logs)", "sample": sample,
                   "window_bits": c_ref, "windows": windows, "setup_s": round(t_setup, 2)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "result_checked": bool(ok)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "result_checked": bool(ok),
    }
    print(json.dumps(line), flush=True)
    if not ok:
        sys.exit(3)


# -------------------------------------------------------------------------------- our arm
def run_ours(args):
    # Everything that libraries print to stdout (e.g. NCCL's version banner) goes to stderr;
    # the real stdout is restored only for the one JSON line.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist

    import bellman_mpc_b200 as bm
    from bellman_mpc_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    w = bm.Worker(local)
    lib = w._lib
    # a non-default stream: the library treats a NULL stream as "use the context's own stream"
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream

    n_total = 1 << args.log_n
    n = n_total // world
    grp = bm.G2 if args.group == "g2" else bm.G1
    pt_bytes = 96 if grp == bm.G1 else 192
    bytes_per_point = pt_bytes + 32                       # affine base + scalar
    mul_per_add = 10 * (1 if grp == bm.G1 else 3)         # 8M+2S; an Fp2 product = 3 Fp products
    g2_gen = bytes.fromhex(
        "13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e"
        "024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8"
        "0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be"
        "0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801")
    curves_gen = bytes.fromhex(
        "17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb"
        "08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1")
    ks = rand_limbs(n, 2 + 7919 * rank)
    t_setup = time.perf_counter()
    bases = bm.Bases.fixed_base_mul(w, grp, curves_gen if grp == bm.G1 else g2_gen, ks)   # resident CRS slice, k_i * G
    if not args.no_precompute:
        bases.precompute()                                               # window tables in HBM (setup)
    scalars_h = torch.from_numpy(rand_limbs(n, 1 + 7919 * rank).view(np.int64)).pin_memory()
    scalars_d = scalars_h.to(dev)
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t_setup

    # a rank's record: its XYZZ partial + its raw flag word (bmpc_multiexp_shard_enqueue_dev)
    rbytes = int(lib.bmpc_shard_record_bytes(grp))
    partial = torch.zeros(rbytes, dtype=torch.uint8, device=dev)
    gathered = torch.zeros(world * rbytes, dtype=torch.uint8, device=dev)
    import ctypes as C
    flags_or = C.c_uint32(0)
    out = np.zeros(pt_bytes, dtype=np.uint8)
    optr = out.ctypes.data_as(C.c_void_p)

    def step_resident(sc_ptr):
        if world == 1:
            st = lib.bmpc_multiexp_dev(w.ctx, bases.handle, 0, sc_ptr, n, None, 0, optr, stream)
        else:
            # shard enqueued, records all-gathered, folded: ONE host synchronisation per step; the status is
            # the reference's for the whole vector (flag words of all ranks ORed, multiexp.rs:244-249)
            st = lib.bmpc_multiexp_shard_enqueue_dev(w.ctx, bases.handle, 0, sc_ptr, n, None, 0, n_total,
                                                     partial.data_ptr(), stream)
            assert st == 0, (st, lib.bmpc_last_error(w.ctx))
            dist.all_gather_into_tensor(gathered, partial)
            st = lib.bmpc_fold_shard_records(w.ctx, grp, gathered.data_ptr(), world, rbytes, optr,
                                             C.byref(flags_or), stream)
            if st == 0:
                st = lib.bmpc_msm_flags_status(flags_or.value)
        assert st == 0, (st, lib.bmpc_last_error(w.ctx))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    # ---- device-resident timing (value) with per-kernel events for the roofline
    for _ in range(args.warmup):
        step_resident(scalars_d.data_ptr())
    lib.bmpc_ctx_profile(w.ctx, 1)
    launches0 = w.launch_count()
    with ClockSampler(local) as clocks:
        ms = timed(lambda: step_resident(scalars_d.data_ptr()), args.steps)
    launches = w.launch_count() - launches0
    prof = {}
    for name, pid in (("accumulate", 0), ("sort", 2), ("reduce", 3)):
        t_ms, cnt = C.c_double(), C.c_uint64()
        lib.bmpc_ctx_profile_read(w.ctx, pid, C.byref(t_ms), C.byref(cnt))
        prof[name] = (t_ms.value / max(cnt.value, 1), int(cnt.value))
    lib.bmpc_ctx_profile(w.ctx, 0)
    result_resident = out.tobytes()
    value = n_total / (ms * 1e-3) / 1e6

    # ---- end to end through the host-buffer call
    def step_e2e():
        if world == 1:
            st = lib.bmpc_multiexp(w.ctx, bases.handle, 0, scalars_h.data_ptr(), n, None, 0, optr)
            assert st == 0, st
        else:
            scalars_d.copy_(scalars_h, non_blocking=True)
            step_resident(scalars_d.data_ptr())

    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_ms_dev = timed(step_e2e, args.steps)
    e2e_wall = (time.perf_counter() - t0) * 1e3 / args.steps
    e2e_serial_ms = max(e2e_ms_dev, 0.0) if world > 1 else e2e_wall   # host call is synchronous: wall == device + copies
    if world > 1:
        t = torch.tensor([e2e_wall], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_serial_ms = float(t.item())
    assert out.tobytes() == result_resident

    # ---- the same, the way the reference's prover issues multiexps (multicore.rs:33-118, prover.rs:233-307:
    # a Waiter per multiexp, the next one issued before the previous is awaited): every step still
    # uploads its scalars from pinned host memory and reads its result back inside the timed region,
    # but the upload of step k+1 overlaps the kernels of step k
    outs = [np.zeros(pt_bytes, dtype=np.uint8) for _ in range(2)]

    def run_pipelined(steps):
        if world == 1:
            pend = []
            for k in range(steps):
                h = C.c_void_p()
                st = lib.bmpc_multiexp_async(w.ctx, bases.handle, 0, scalars_h.data_ptr(), n, None, 0, C.byref(h))
                assert st == 0, (st, lib.bmpc_last_error(w.ctx))
                pend.append((h, outs[k % 2]))
                if len(pend) == 2:
                    hh, oo = pend.pop(0)
                    assert lib.bmpc_waiter_wait(hh, oo.ctypes.data_as(C.c_void_p)) == 0
            for hh, oo in pend:
                assert lib.bmpc_waiter_wait(hh, oo.ctypes.data_as(C.c_void_p)) == 0
            return
        bufs = [scalars_d, scalars_d2]
        evs = [None, None]

        def issue(k):
            with torch.cuda.stream(copy_stream):
                bufs[k % 2].copy_(scalars_h, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                evs[k % 2] = ev
        issue(0)
        for k in range(steps):
            if k + 1 < steps:
                issue(k + 1)        # its buffer was last read by step k-1, which has completed (host-blocking call)
            tstream.wait_event(evs[k % 2])
            step_resident(bufs[k % 2].data_ptr())

    if world > 1:
        scalars_d2 = torch.empty_like(scalars_d)
        copy_stream = torch.cuda.Stream(device=dev)
    run_pipelined(2)
    barrier()
    t0 = time.perf_counter()
    run_pipelined(args.steps)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    barrier()
    e2e_checked = bool((outs[0].tobytes() == result_resident and outs[1].tobytes() == result_resident)
                       if world == 1 else out.tobytes() == result_resident)
    assert e2e_checked

    # ---- the headline result itself: (sum over all ranks of sum_i k_i s_i) G, no MSM involved
    result_checked = None
    if rank == 0:
        from oracle import cref as _cref, fields as _fields
        tot = 0
        for r in range(world):
            ks_r = ks if r == 0 else rand_limbs(n, 2 + 7919 * r)
            sc_r = scalars_h.numpy().view(np.uint64) if r == 0 else rand_limbs(n, 1 + 7919 * r)
            tot += limbs_to_int(_cref.fr_dot(ks_r, sc_r))
        tot %= _fields.Fr.p
        if grp == bm.G1:
            expect = _cref.g1_generator_mul(int_to_limbs(tot))
        else:
            from oracle import curves as _curves
            expect = _curves.G2.to_uncompressed(_curves.G2.mul(_curves.G2.gen, tot))
        result_checked = bool(expect == result_resident)

    peaks = measured_peaks()
    acc_ms, acc_cnt = prof["accumulate"]
    cw, ww, hh = C.c_uint32(), C.c_uint32(), C.c_uint32()
    lib.bmpc_msm_geometry(w.ctx, bases.handle, n, C.byref(cw), C.byref(ww), C.byref(hh))
    ainfo = (C.c_uint32 * 8)()
    lib.bmpc_msm_accumulate_info(w.ctx, bases.handle, n, C.byref(ainfo))
    acc_mode = int(ainfo[0])
    kname = {0: "msm_accumulate_kernel", 1: "msm_accumulate_affine_kernel", 2: "msm_pair_round_kernel"}[acc_mode]
    kname += "<%s>" % ("Fp" if grp == bm.G1 else "Fp2")
    headline_cfg = world == 1 and args.log_n == 24 and grp == bm.G1 and not args.no_precompute
    traffic = None
    if headline_cfg:
        traffic = {0: peaks.get("acc_traffic_bytes"), 1: peaks.get("aff_traffic_bytes"),
                   2: peaks.get("pair_traffic_bytes")}[acc_mode]
    # The bucket accumulation is bound by the 32-bit multiplier pipe, so THAT is the roofline:
    # field products of the point additions actually issued (W windows per point; 10 per XYZZ mixed
    # addition, 6 per batched-affine addition: 1 running product, 2 to unwind it, slope, slope^2, y3;
    # the shared inversion and its product tree are overhead, not counted) x 300 MAC32 per Fp product,
    # against the measured mad.lo.cc/madc.hi.cc rate of this GPU.
    mul_per_add = (10 if acc_mode == 0 else 6) * (1 if grp == bm.G1 else 3)
    roofline = None
    if peaks.get("mac32_per_s") and acc_ms:
        executed = mul_per_add * 300 * ww.value * n
        ach = executed / (acc_ms * 1e-3)
        roofline = {"kernel": kname, "bound": "int32-mac", "achieved": ach / 1e12, "peak": peaks["mac32_per_s"] / 1e12,
                    "unit": "TMAC32/s", "frac": ach / peaks["mac32_per_s"],
                    "traffic": traffic, "algorithmic_bytes": bytes_per_point * n,
                    "traffic_ratio": (traffic / (bytes_per_point * n)) if traffic else None,
                    "traffic_source": "profiles/r02_ncu_kernel_summaries.json (ncu --set full of this workload, 1 GPU, 2^24): dram__bytes_read.sum + dram__bytes_write.sum per launch",
                    "peak_source": "profiles/r01_imad_peak.json (bench/imad_peak.cu on this pool's B200: 32 MAC32/clk/SM)",
                    "work": f"executed: {ww.value} windows of c={cw.value} bits x {mul_per_add} Fp products x 300 MAC32 per point",
                    "kernel_ms": acc_ms, "launches_per_step": acc_cnt // max(args.steps, 1),
                    "share_of_step": acc_ms * (acc_cnt // max(args.steps, 1)) / ms if ms else None,
                    "ncu_fmaheavy_pct": peaks.get("fmaheavy_pct") if headline_cfg else None,
                    "survey_model_frac": None,
                    "survey_model_note": "SURVEY 8d charges 48000 MAC32 per point (16 windows x 10 products); this path "
                                         "issues fewer products than that model, so a fraction against it would exceed 1 and is not reported"}
    hbm_ach = bytes_per_point * n / (acc_ms * 1e-3) / 1e9 if acc_ms else None
    roofline_hbm = {"kernel": kname, "bound": "hbm", "achieved": hbm_ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": (hbm_ach / peaks["hbm_gbs"]) if hbm_ach else None,
                    "peak_source": f"MEASURED_PEAKS.json ({peaks['source']})",
                    "dram_gbs_measured": (traffic / (acc_ms * 1e-3) / 1e9) if (traffic and acc_ms) else None,
                    "note": "not the binding roofline: algorithmic bytes (128 B per point) over the kernel time"}

    line = {
        "metric": METRIC if (grp == bm.G1 and args.log_n == 24) else f"{args.group}_msm_mpts_per_s_2p{args.log_n}",
        "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u32-limb Montgomery (Fp 12x32, Fr 8x32)", "data": "synthetic",
        "config": {"workload": f"{args.group.upper()} multiexp 2^{args.log_n} points, uniform 254-bit scalars, FullDensity "
                               f"(BASELINE configs[1] shape at the size the metric is quoted on)",
                   "bases": "k_i*G, resident (CRS registered once%s)" % ("" if args.no_precompute else
                            ", window tables 2^(cw)*P_i precomputed at registration"), "points_per_gpu": n,
                   "l2": "inputs larger than L2 (scalars %d MiB + bases %d MiB per GPU%s)" % (
                       n * 32 >> 20, n * pt_bytes >> 20,
                       "" if args.no_precompute else "; with the %d window tables %d MiB per GPU" % (ww.value, ww.value * n * pt_bytes >> 20)),
                   "table_bytes_per_gpu": 0 if args.no_precompute else ww.value * n * pt_bytes,
                   "parallelism": f"bases split x{world}, all-gather of {rbytes}-byte records (XYZZ partial + flag word), one host synchronisation per step" if world > 1 else "single GPU",
                   "setup_s": round(t_setup, 2)},
        "clocks": clocks.summary(),
        "e2e": {"value": n_total / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": 96 + 64,
                "how": "host-buffer multiexps issued as the reference's prover issues them (a Waiter each, two in "
                       "flight: bmpc_multiexp_async / bmpc_waiter_wait" + ("" if world == 1 else "; per rank a second scalar "
                       "buffer filled on a copy stream") + "): every step uploads its scalars from pinned host memory "
                       "and reads its result back inside the timed region; wall clock over the steps",
                "serial_ms_per_step": e2e_serial_ms, "serial_value": n_total / (e2e_serial_ms * 1e-3) / 1e6,
                "result_bytes_equal_resident": e2e_checked},
        "gpu_launches": launches, "result_checked": result_checked,
        "roofline": roofline, "roofline_hbm": roofline_hbm,
        "kernel_ms": {k: v[0] for k, v in prof.items()},
    }

    # ---- CPU baseline beside it (rank 0, N = 1): the oracle's C port on a bounded sample of the same
    # vector, with the window the reference picks for the WHOLE vector (c = 17 at 2^24), + parity check
    if world == 1 and not args.no_cpu_baseline and grp == bm.G1:
        from oracle import cref
        threads = cref.hardware_threads()
        slog = min(args.cpu_sample_log, args.log_n)
        ns = 1 << slog
        cb = cref.CBases.from_uncompressed(1, bases.read(0, ns))
        sc = np.ascontiguousarray(scalars_h.numpy().view(np.uint64)[:ns])
        t0 = time.perf_counter()
        st, cpu_out = cref.multiexp_window(cb, 0, sc, n, threads=threads)
        cpu_s = time.perf_counter() - t0
        gpu_out = bm.multiexp(w, (bases, 0), bm.FullDensity(), sc).wait()
        c_ref = int(cref.load().orc_window_size(n))
        line["cpu_baseline"] = {"value": ns / cpu_s / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"first 2^{slog} points of the workload, reference window c={c_ref} "
                                          f"({(255 + c_ref - 1) // c_ref} windows) as for the whole 2^{args.log_n} vector",
                                "seconds": cpu_s, "matches_gpu_bytes": bool(st == 0 and cpu_out == gpu_out)}
        cb.free()

    # ---- the same multiexp over bases WITHOUT window tables (a base vector that is not a resident CRS)
    if world == 1 and not args.no_precompute and not args.no_plain and grp == bm.G1:
        try:
            plain = bm.Bases.fixed_base_mul(w, grp, curves_gen, ks)
            for _ in range(2):
                st = lib.bmpc_multiexp_dev(w.ctx, plain.handle, 0, scalars_d.data_ptr(), n, None, 0, optr, stream)
                assert st == 0
            same = out.tobytes() == result_resident
            bases_keep, bases = bases, plain
            ms_plain = timed(lambda: step_resident(scalars_d.data_ptr()), max(2, args.steps // 2))
            bases = bases_keep
            cp, wp, hp = C.c_uint32(), C.c_uint32(), C.c_uint32()
            lib.bmpc_msm_geometry(w.ctx, plain.handle, n, C.byref(cp), C.byref(wp), C.byref(hp))
            line["no_precompute"] = {"value": n_total / (ms_plain * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms_plain,
                                     "window_bits": cp.value, "windows": wp.value, "bucket_sets": hp.value,
                                     "same_result_bytes": bool(same)}
            plain.free()
        except Exception as e:
            line["no_precompute"] = {"error": repr(e)}

    # ---- Groth16 prove @ 2^prove_log_n (N = 1), same run
    if world == 1 and not args.no_prove and grp == bm.G1:
        try:
            from bench_prove import prove_bench
            line["prove"] = prove_bench(w, args.prove_log_n, max(2, args.steps // 2), args.no_cpu_baseline,
                                        precompute=not args.no_precompute)
        except Exception as e:  # keep the headline line even if the secondary workload fails
            line["prove"] = {"error": repr(e)}

    if world > 1 and not args.no_prove and grp == bm.G1:
        try:
            from bench_prove import prove_bench_sharded
            line["prove"] = prove_bench_sharded(w, args.prove_log_n, max(2, args.steps // 2), world, rank,
                                                precompute=not args.no_precompute)
        except Exception as e:
            line["prove"] = {"error": repr(e)}

    # ---- constraint evaluation + key generation on the device
    if world == 1 and not args.no_r1cs and grp == bm.G1:
        try:
            sys.path.insert(0, os.path.join(ROOT, "bench"))
            import r1cs_bench
            line["r1cs"] = r1cs_bench.run(w, args.r1cs_log_n)
        except Exception as e:
            line["r1cs"] = {"error": repr(e)}

    # ---- EvaluationDomain sweep (BASELINE config #3): fft on resident coefficients, GB/s vs HBM peak
    if world == 1 and not args.no_ntt:
        ntt = []
        for logm in range(16, args.ntt_max_log + 1, 2):
            m = 1 << logm
            coeffs = torch.from_numpy(rand_limbs(m, 3).view(np.int64)).to(dev)
            lib.bmpc_ntt_dev(w.ctx, coeffs.data_ptr(), logm, 0, stream)
            torch.cuda.synchronize()
            lib.bmpc_ctx_profile(w.ctx, 1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record()
            for _ in range(reps):
                st = lib.bmpc_ntt_dev(w.ctx, coeffs.data_ptr(), logm, 0, stream)
                assert st == 0
            e1.record()
            torch.cuda.synchronize()
            t_ms, cnt = C.c_double(), C.c_uint64()
            lib.bmpc_ctx_profile_read(w.ctx, 1, C.byref(t_ms), C.byref(cnt))
            lib.bmpc_ctx_profile(w.ctx, 0)
            ms_t = e0.elapsed_time(e1) / reps
            # the same transform through the host-buffer call (what a drop-in `fft(&worker)` does): pinned
            # coefficients -> H2D -> passes -> D2H, wall clock
            e2e_ntt = None
            if logm <= 24:
                hbuf = torch.from_numpy(rand_limbs(m, 3).view(np.int64)).pin_memory()
                lib.bmpc_ntt(w.ctx, hbuf.data_ptr(), logm, 0)
                t0 = time.perf_counter()
                for _ in range(3):
                    assert lib.bmpc_ntt(w.ctx, hbuf.data_ptr(), logm, 0) == 0
                e2e_ntt = (time.perf_counter() - t0) * 1e3 / 3
                del hbuf
            passes = cnt.value // reps
            pass_ms = t_ms.value / max(cnt.value, 1)
            butterflies = (m // 2) * logm
            ntt.append({"log_m": logm, "ms": ms_t, "passes": int(passes), "e2e_host_buffer_ms": e2e_ntt,
                        "algorithmic_gbs": 64 * m / (ms_t * 1e-3) / 1e9,
                        "hbm_frac": 64 * m / (ms_t * 1e-3) / 1e9 / peaks["hbm_gbs"],
                        "pass_kernel_gbs": 64 * m / (pass_ms * 1e-3) / 1e9,
                        "tmac32_per_s": butterflies * 136 / (ms_t * 1e-3) / 1e12,
                        "int_frac": (butterflies * 136 / (ms_t * 1e-3) / peaks["mac32_per_s"]) if peaks.get("mac32_per_s") else None})
            del coeffs
        line["ntt"] = {"op": "EvaluationDomain::fft, coefficients resident, in place", "sweep": ntt,
                       "bytes_per_coeff": 64, "mac32_per_butterfly": 136}

    # ---- distributed EvaluationDomain::fft over the N GPUs (four-step, all-to-all over NCCL), N > 1
    if world > 1 and not args.no_ntt and (world & (world - 1)) == 0:
        try:
            from bellman_mpc_b200 import dist as bdist
            logm = args.ntt_dist_log
            plan = bdist.FourStepPlan(logm, world)
            ops, a2a = bdist.GpuFrOps(w), bdist.torch_all_to_all()
            local0 = torch.from_numpy(rand_limbs(plan.local, 30 + rank).view(np.int64)).to(dev)
            for _ in range(2):
                out = bdist.distributed_transform(local0.clone(), plan, rank, bm.FFT, ops, a2a)
            # the inverse transform of the result must give the input back on every rank
            back = bdist.distributed_transform(out.clone(), plan, rank, bm.IFFT, ops, a2a)
            ok = torch.tensor([int(torch.equal(back, local0))], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            reps = 5
            ins = [local0.clone() for _ in range(reps)]
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for r in range(reps):
                bdist.distributed_transform(ins[r], plan, rank, bm.FFT, ops, a2a)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_t = float(t.item())
            m = 1 << logm
            line["ntt_dist"] = {"op": "EvaluationDomain::fft over a domain split across the GPUs (four-step, 3 all-to-all)",
                                "log_m": logm, "n_gpus": world, "ms": ms_t, "roundtrip_identical": bool(ok.item()),
                                "algorithmic_gbs": 64 * m / (ms_t * 1e-3) / 1e9,
                                "tmac32_per_s": (m // 2) * logm * 136 / (ms_t * 1e-3) / 1e12,
                                "all_to_all_bytes_per_gpu": 3 * plan.local * 32 * (world - 1) // world}
            del ins, local0, out, back
        except Exception as e:
            line["ntt_dist"] = {"error": repr(e)}

    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)
    bases.free()
    w.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=24)
    ap.add_argument("--group", default="g1", choices=["g1", "g2"])
    ap.add_argument("--ref-sample-log", type=int, default=23,
                    help="reference arm: points per step (a sample of the 2^log_n vector, window as for the whole vector)")
    ap.add_argument("--cpu-sample-log", type=int, default=22, help="cpu_baseline leg of our arm: sample size")
    ap.add_argument("--no-plain", action="store_true", help="skip the table-free timing of the same multiexp")
    ap.add_argument("--prove-log-n", type=int, default=22)
    ap.add_argument("--no-prove", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-precompute", action="store_true")
    ap.add_argument("--no-ntt", action="store_true")
    ap.add_argument("--no-r1cs", action="store_true")
    ap.add_argument("--r1cs-log-n", type=int, default=18)
    ap.add_argument("--ntt-max-log", type=int, default=26)
    ap.add_argument("--ntt-dist-log", type=int, default=26)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
