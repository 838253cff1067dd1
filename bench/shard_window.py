"""Window-size sweep for a shard-sized (2^k-point) G1 multiexp over table-registered bases:
python bench/shard_window.py [log_n] [c ...]  -> ms per multiexp and the per-phase split for each c."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bellman_mpc_b200 as bm  # noqa: E402
from bench import rand_limbs  # noqa: E402

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 21
cs = [int(x) for x in sys.argv[2:]] or [18, 19, 20]
n = 1 << log_n
w = bm.Worker(0)
lib = w._lib
ts = torch.cuda.Stream()
torch.cuda.set_stream(ts)
gen = bytes.fromhex(
    "17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb"
    "08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1")
ks = rand_limbs(n, 2)
sc = torch.from_numpy(rand_limbs(n, 1).view(np.int64)).cuda()
out = np.zeros(96, dtype=np.uint8)
optr = out.ctypes.data_as(C.c_void_p)
ref = None
for c in cs:
    bases = bm.Bases.fixed_base_mul(w, bm.G1, gen, ks).precompute(c)
    for _ in range(3):
        assert lib.bmpc_multiexp_dev(w.ctx, bases.handle, 0, sc.data_ptr(), n, None, 0, optr, ts.cuda_stream) == 0
    lib.bmpc_ctx_profile(w.ctx, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        assert lib.bmpc_multiexp_dev(w.ctx, bases.handle, 0, sc.data_ptr(), n, None, 0, optr, ts.cuda_stream) == 0
    e1.record()
    torch.cuda.synchronize()
    prof = {}
    for name, pid in (("accumulate", 0), ("sort", 2), ("reduce", 3)):
        t_ms, cnt = C.c_double(), C.c_uint64()
        lib.bmpc_ctx_profile_read(w.ctx, pid, C.byref(t_ms), C.byref(cnt))
        prof[name] = round(t_ms.value / max(cnt.value, 1), 3)
    lib.bmpc_ctx_profile(w.ctx, 0)
    if ref is None:
        ref = out.tobytes()
    assert out.tobytes() == ref, "result depends on the window size"
    print("c=%d  %.3f ms  %s" % (c, e0.elapsed_time(e1) / 5, prof), flush=True)
    bases.free()
