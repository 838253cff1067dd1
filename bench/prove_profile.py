"""Two proofs of the synthetic 2^k workload, for `ncu` launch lists (profiles/): 
python bench/prove_profile.py [log_m]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bellman_mpc_b200 as bm  # noqa: E402
from bench_prove import Workload  # noqa: E402

log_m = int(sys.argv[1]) if len(sys.argv) > 1 else 22
w = bm.Worker(0)
wl = Workload(w, log_m, precompute=True)
wl.prove()
t0 = time.perf_counter()
proof = wl.prove()
print("prove_s", time.perf_counter() - t0, proof[:8].hex())
