"""BASELINE config #5, the part beside the sharded multiexps (those are `bench.py --log-n 26 --group g1|g2`
under torchrun): the ceremony's batch scalar multiplication (groth16/mpc.rs:647-706 make_new_paramter /
make_new_tau_paramter) at 2^22 points in G1 and G2, per-element scalars and one shared scalar, the
points split over the ranks (embarrassingly parallel, no collective).  One JSON line per case on rank 0.

  python bench/config5.py [--log-n 22]                      (1 GPU)
  python -m torch.distributed.run --nproc-per-node N ... bench/config5.py --gpus N
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bellman_mpc_b200 as bm  # noqa: E402
from bench import rand_limbs, limbs_to_int, measured_peaks  # noqa: E402
import bench_prove  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--log-n", type=int, default=22)
    ap.add_argument("--steps", type=int, default=2)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    w = bm.Worker(local)
    n_total = 1 << args.log_n
    n = n_total // world
    peaks = measured_peaks()
    Q = bm.FR_MODULUS

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for grp, name, gen in ((bm.G1, "g1", bench_prove.G1_GEN), (bm.G2, "g2", bench_prove.G2_GEN)):
        ks = rand_limbs(n, 50 + rank)
        bases = bm.Bases.fixed_base_mul(w, grp, gen, ks)
        for per_element in (True, False):
            sc = rand_limbs(n if per_element else 1, 60 + rank * 2 + int(per_element))
            out = bases.scalar_mul(sc, per_element=per_element)       # warm-up
            out.free()
            times = []
            for _ in range(args.steps):
                barrier()
                t0 = time.perf_counter()
                out = bases.scalar_mul(sc, per_element=per_element)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                if world > 1:
                    t = torch.tensor([dt], device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    dt = float(t.item())
                times.append(dt)
                if _ + 1 < args.steps:
                    out.free()
            # known discrete logs: out[i] = (k_i s_i) G, checked on a sample through the fixed-base path
            idx = np.random.RandomState(7).randint(0, n, size=256)
            prod = [limbs_to_int(ks[i]) * limbs_to_int(sc[i if per_element else 0]) % Q for i in idx]
            want = bm.Bases.fixed_base_mul(w, grp, gen, bm.ints_to_limbs(prod))
            pb = 96 if grp == bm.G1 else 192
            got = b"".join(out.read(int(i), 1) for i in idx[:64])
            ok = got == want.read(0, 64)
            want.free()
            out.free()
            ok_t = torch.tensor([int(ok)], device=dev)
            if world > 1:
                dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
            best = min(times)
            # executed work: double-and-add over the scalar's bits: 255 doublings (9 field products) + one
            # mixed addition (10) per set bit (~127), an Fp2 product = 3 Fp products, 300 MAC32 each
            fp_per = 3 if grp == bm.G2 else 1
            mac = (255 * 9 + 127 * 10) * 300 * fp_per * n_total
            line = {"metric": f"batch_scalar_mul_{name}_{'per_element' if per_element else 'same_scalar'}_mpts_per_s",
                    "value": n_total / best / 1e6, "unit": "Mpts/s", "n_gpus": world, "points": n_total,
                    "seconds": best, "all_s": [round(t, 4) for t in times],
                    "timed_region": "bmpc_batch_scalar_mul: host scalars -> H2D -> double-and-add per point -> canonical affine, resident output",
                    "sample_matches_known_dlog": bool(ok_t.item()),
                    "roofline": {"bound": "int32-mac", "achieved": mac / best / 1e12,
                                 "peak": (peaks.get("mac32_per_s") or 0) * world / 1e12, "unit": "TMAC32/s",
                                 "frac": (mac / best / (peaks["mac32_per_s"] * world)) if peaks.get("mac32_per_s") else None,
                                 "work": "255 doublings x 9 + ~127 mixed additions x 10 field products per point"},
                    "reference": "groth16/mpc.rs:647-706"}
            if rank == 0:
                print(json.dumps(line), flush=True)
        bases.free()
    w.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
