"""Same dense points, two ways: a FullDensity multiexp over n_dense scalars against a density-mapped one
over 2 n_dense positions (half of them dense) -- what create_proof's B multiexps look like.
  python bench/density_ab.py [--group g2] [--log-dense 21]"""
import argparse, ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bellman_mpc_b200 as bm
from bench import rand_limbs
import bench_prove

ap = argparse.ArgumentParser()
ap.add_argument("--group", default="g2")
ap.add_argument("--log-dense", type=int, default=21)
args = ap.parse_args()
w = bm.Worker(0)
lib = w._lib
grp = bm.G2 if args.group == "g2" else bm.G1
gen = bench_prove.G2_GEN if grp == bm.G2 else bench_prove.G1_GEN
nd = 1 << args.log_dense
bases = bm.Bases.fixed_base_mul(w, grp, gen, rand_limbs(nd, 2)).precompute()
rs = np.random.RandomState(5)
npos = 2 * nd
bits = np.zeros(npos, dtype=bool)
bits[rs.permutation(npos)[:nd]] = True
dens = bm.DensityTracker.from_bits(bits)
sc_pos = rand_limbs(npos, 3)
sc_dense = np.ascontiguousarray(sc_pos[bits])
out1, out2 = np.zeros(192, dtype=np.uint8), np.zeros(192, dtype=np.uint8)
st = torch.cuda.Stream()
d_pos = torch.from_numpy(sc_pos.view(np.int64)).cuda()
d_dense = torch.from_numpy(sc_dense.view(np.int64)).cuda()
d_words = torch.from_numpy(dens.words().view(np.int64)).cuda()


def run(kind, hint):
    o = out1 if kind == "full" else out2
    if kind == "full":
        f = lambda: lib.bmpc_multiexp_dev(w.ctx, bases.handle, 0, d_dense.data_ptr(), nd, None, 0, o.ctypes.data_as(C.c_void_p), st.cuda_stream)
    elif hint:
        f = lambda: lib.bmpc_multiexp(w.ctx, bases.handle, 0, sc_pos.ctypes.data_as(C.c_void_p), npos, dens.words().ctypes.data_as(C.c_void_p), npos, o.ctypes.data_as(C.c_void_p))
    else:
        f = lambda: lib.bmpc_multiexp_dev(w.ctx, bases.handle, 0, d_pos.data_ptr(), npos, d_words.data_ptr(), npos, o.ctypes.data_as(C.c_void_p), st.cuda_stream)
    for _ in range(2):
        assert f() == 0
    lib.bmpc_ctx_profile(w.ctx, 1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / 3
    prof = {}
    for name, pid in (("accumulate", 0), ("sort", 2), ("reduce", 3)):
        t_ms, cnt = C.c_double(), C.c_uint64()
        lib.bmpc_ctx_profile_read(w.ctx, pid, C.byref(t_ms), C.byref(cnt))
        prof[name] = round(t_ms.value / 3, 3)
    lib.bmpc_ctx_profile(w.ctx, 0)
    info = (C.c_uint32 * 8)()
    print(json.dumps({"kind": kind, "hint": hint, "ms": round(ms, 3), **prof}), flush=True)


run("full", False)
run("density", False)
run("density", True)
print("same result", out1.tobytes() == out2.tobytes())
