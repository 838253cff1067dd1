"""A/B timing of create_proof at 2^k under environment knobs, same process-independent workload:
python bench/prove_ab.py [log_m] [reps]  -> prints min / median seconds"""
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bellman_mpc_b200 as bm  # noqa: E402
from bench_prove import Workload  # noqa: E402

log_m = int(sys.argv[1]) if len(sys.argv) > 1 else 22
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
w = bm.Worker(0)
wl = Workload(w, log_m, precompute=True)
wl.prove()
wl.prove()
ts = []
for _ in range(reps):
    t0 = time.perf_counter()
    proof = wl.prove()
    ts.append(time.perf_counter() - t0)
print("prove_s min %.4f median %.4f" % (min(ts), statistics.median(ts)), proof[:8].hex(), flush=True)
