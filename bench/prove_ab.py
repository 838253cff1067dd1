"""A/B timing of create_proof at 2^k under environment knobs, same workload, one process:
  python bench/prove_ab.py [log_m] [reps] [NAME=v1,v2,...] [world]
prints min / median seconds per value of the swept BMPC_* variable and whether the proof equals the
known-dlog expectation.  world > 1: times ONE rank's share of the sharded proof (rank world/2:
bmpc_create_proof_partials over its slices) -- what an N-GPU run costs per rank, measured on one GPU."""
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bellman_mpc_b200 as bm  # noqa: E402
from bellman_mpc_b200 import dist as bdist  # noqa: E402
from bench_prove import Workload  # noqa: E402

log_m = int(sys.argv[1]) if len(sys.argv) > 1 else 22
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
sweep = sys.argv[3] if len(sys.argv) > 3 else ""
world = int(sys.argv[4]) if len(sys.argv) > 4 else 1
name, vals = (sweep.split("=")[0], sweep.split("=")[1].split(",")) if sweep else (None, [None])
w = bm.Worker(0)
wl = Workload(w, log_m, precompute=True, world=world, rank=world // 2)
expect = wl.expected_proof() if world == 1 else None


def run():
    if world == 1:
        return wl.prove()
    pb, st = bdist.proof_partials(w, wl.params, wl.assignment, wl.plan)
    assert pb is not None, st
    return pb


for v in vals:
    if name:
        os.environ[name] = v
        w.reload_env()
    run()
    run()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        proof = run()
        ts.append(time.perf_counter() - t0)
    print(json.dumps({"log_m": log_m, "world": world, "sweep": f"{name}={v}" if name else None,
                      "min_s": round(min(ts), 4), "median_s": round(statistics.median(ts), 4),
                      "matches_expectation": (proof == expect) if world == 1 else None}), flush=True)
