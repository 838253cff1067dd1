"""ONE process, all GPUs of the box behind one call (SURVEY 8b `bmpc_ctx_create(devices, n)`;
csrc/multi.cu): a 2^log_n G1 multiexp through bmpc_multi_multiexp (host scalars in, 96 bytes out) and a
2^log_m proof through bmpc_multi_create_proof, timed by wall clock, results checked against the
known-dlog expectation and the single-device bytes.
  python bench/multi_onecall.py [--devices 8] [--log-n 24] [--log-m 22]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bellman_mpc_b200 as bm  # noqa: E402
from bench import rand_limbs, limbs_to_int, int_to_limbs  # noqa: E402
import bench_prove  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--devices", type=int, default=0)
    ap.add_argument("--log-n", type=int, default=24)
    ap.add_argument("--log-m", type=int, default=22)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    nd = args.devices or torch.cuda.device_count()
    mw = bm.MultiWorker(list(range(nd)))
    lib = mw._lib
    n = 1 << args.log_n
    # bases with known dlogs, made slice by slice on each device's own context, then registered split
    ks = rand_limbs(n, 2)
    w0 = bm.Worker(0)
    t0 = time.perf_counter()
    raw = bm.Bases.fixed_base_mul(w0, bm.G1, bench_prove.G1_GEN, ks)
    blob = np.frombuffer(raw.read(), dtype=np.uint8)
    raw.free()
    mb = bm.MultiBases.from_uncompressed(mw, bm.G1, blob)
    mb.precompute()
    setup_s = time.perf_counter() - t0
    sc = torch.from_numpy(rand_limbs(n, 1).view(np.int64)).pin_memory()
    out = np.zeros(96, dtype=np.uint8)
    from oracle import cref, fields
    tot = limbs_to_int(cref.fr_dot(ks, sc.numpy().view(np.uint64))) % fields.Fr.p
    expect = cref.g1_generator_mul(int_to_limbs(tot))

    def step():
        rc = lib.bmpc_multi_multiexp(mw.handle, mb.handle, 0, sc.data_ptr(), n, None, 0, out.ctypes.data_as(C.c_void_p))
        assert rc == 0, (rc, mw.last_error())
    for _ in range(3):
        step()
    ok = out.tobytes() == expect
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    ms = (time.perf_counter() - t0) * 1e3 / args.steps
    print(json.dumps({"metric": "g1_msm_mpts_per_s_one_call", "devices": nd, "log_n": args.log_n,
                      "value": n / ms / 1e3, "unit": "Mpts/s", "ms_per_call": ms,
                      "timed_region": "bmpc_multi_multiexp: pinned host scalars -> per-device H2D of its slice -> shard "
                                      "multiexps on host threads -> peer gather of the XYZZ partials on device 0 -> fold -> 96 B",
                      "result_checked": bool(ok), "setup_s": round(setup_s, 2)}), flush=True)
    mb.free()
    # ---- create_proof through one call
    wl = bench_prove.Workload(w0, args.log_m)
    expect_proof = wl.expected_proof()
    split = lambda b: bm.MultiBases.from_uncompressed(mw, b.group, np.frombuffer(b.read(), dtype=np.uint8)).precompute()
    p = wl.params
    mp = bm.MultiParameters(mw, split(wl.h), split(wl.l), split(wl.qa), split(wl.qb1), split(wl.qb2),
                            p.alpha_g1, p.beta_g1, p.beta_g2, p.delta_g1, p.delta_g2)
    wl.free()
    for _ in range(2):
        proof = bm.create_proof(wl.assignment, mp, wl.r, wl.s)
    ts = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        proof = bm.create_proof(wl.assignment, mp, wl.r, wl.s)
        ts.append(time.perf_counter() - t0)
    print(json.dumps({"metric": "groth16_prove_seconds_one_call", "devices": nd, "log_m": args.log_m,
                      "value": min(ts), "unit": "s", "all_s": [round(t, 4) for t in ts],
                      "timed_region": "bmpc_multi_create_proof: pinned host a, b, c + assignments -> devices 0-2 upload and transform "
                                      "one of a, b, c each, device 0 combines them into H, every device pulls its slice; every "
                                      "device: H2D of its assignment slices, its share of the eight multiexps -> partial sums "
                                      "folded and tail on device 0",
                      "matches_known_dlog_expectation": bool(proof == expect_proof)}), flush=True)
    mp.free()
    mw.close()
    w0.close()


if __name__ == "__main__":
    main()
