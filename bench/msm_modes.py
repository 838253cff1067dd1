"""Times the G1 (or G2) multiexp at 2^log_n under different accumulate kernels / knobs (env), one
process, resident inputs, known-dlog result check.  Scratch harness for tuning runs:
  python bench/msm_modes.py --log-n 24 --modes pairs,affine --k 256,384"""
import argparse, ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bellman_mpc_b200 as bm
from bench import rand_limbs, limbs_to_int, int_to_limbs
import bench_prove


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=24)
    ap.add_argument("--group", default="g1")
    ap.add_argument("--modes", default="pairs,affine")
    ap.add_argument("--k", default="256")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--no-precompute", action="store_true")
    ap.add_argument("--sweep", default="", help="NAME=v1,v2,...: run every mode once per value of this BMPC_* variable")
    args = ap.parse_args()
    w = bm.Worker(0)
    lib = w._lib
    n = 1 << args.log_n
    grp = bm.G1 if args.group == "g1" else bm.G2
    gen = bench_prove.G1_GEN if grp == bm.G1 else bench_prove.G2_GEN
    ks = rand_limbs(n, 2)
    bases = bm.Bases.fixed_base_mul(w, grp, gen, ks)
    if not args.no_precompute:
        bases.precompute()
    sc_h = rand_limbs(n, 1)
    sc = torch.from_numpy(sc_h.view(np.int64)).cuda()
    from oracle import cref, fields, curves
    tot = limbs_to_int(cref.fr_dot(ks, sc_h)) % fields.Fr.p
    if grp == bm.G1:
        expect = cref.g1_generator_mul(int_to_limbs(tot))
    else:
        expect = curves.G2.to_uncompressed(curves.G2.mul(curves.G2.gen, tot))
    out = np.zeros(96 if grp == bm.G1 else 192, dtype=np.uint8)
    optr = out.ctypes.data_as(C.c_void_p)
    st = torch.cuda.Stream()
    res = []
    sweep_name, sweep_vals = None, [None]
    if args.sweep:
        sweep_name, vals = args.sweep.split("=")
        sweep_vals = vals.split(",")
    for mode, sv in [(m, v) for m in args.modes.split(",") for v in sweep_vals]:
        if sweep_name:
            os.environ[sweep_name] = sv
        for k in args.k.split(","):
            os.environ["BMPC_ACC_PAIRS"] = "1" if mode == "pairs" else "0"
            os.environ["BMPC_ACC_AFFINE"] = "0" if mode == "xyzz" else "-1"
            if mode == "xyzz":
                os.environ["BMPC_ACC_AFFINE"] = "0"
            else:
                os.environ.pop("BMPC_ACC_AFFINE", None)
            os.environ["BMPC_PAIR_K"] = k
            w.reload_env()
            for _ in range(2):
                rc = lib.bmpc_multiexp_dev(w.ctx, bases.handle, 0, sc.data_ptr(), n, None, 0, optr, st.cuda_stream)
                assert rc == 0, (rc, lib.bmpc_last_error(w.ctx))
            ok = out.tobytes() == expect
            lib.bmpc_ctx_profile(w.ctx, 1)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                lib.bmpc_multiexp_dev(w.ctx, bases.handle, 0, sc.data_ptr(), n, None, 0, optr, st.cuda_stream)
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) * 1e3 / args.steps
            prof = {}
            for name, pid in (("accumulate", 0), ("sort", 2), ("reduce", 3)):
                t_ms, cnt = C.c_double(), C.c_uint64()
                lib.bmpc_ctx_profile_read(w.ctx, pid, C.byref(t_ms), C.byref(cnt))
                prof[name] = round(t_ms.value / max(args.steps, 1), 3)
            lib.bmpc_ctx_profile(w.ctx, 0)
            info = (C.c_uint32 * 8)()
            lib.bmpc_msm_accumulate_info(w.ctx, bases.handle, n, C.byref(info))
            r = {"mode": mode, "sweep": f"{sweep_name}={sv}" if sweep_name else None, "k": int(k), "ms": round(ms, 3), "mpts": round(n / ms / 1e3, 1), "ok": bool(ok),
                 "info": list(info)[:6], **prof}
            print(json.dumps(r), flush=True)
            res.append(r)
            if mode != "pairs":
                break
    bases.free()
    w.close()


if __name__ == "__main__":
    main()
