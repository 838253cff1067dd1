"""A few forward transforms at 2^log_m on resident coefficients, for ncu captures (profiles/):
python bench/ntt_profile.py [log_m] [reps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bellman_mpc_b200 as bm  # noqa: E402
from bench import rand_limbs  # noqa: E402

log_m = int(sys.argv[1]) if len(sys.argv) > 1 else 24
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w = bm.Worker(0)
st = torch.cuda.Stream()
coeffs = torch.from_numpy(rand_limbs(1 << log_m, 3).view(np.int64)).cuda()
for _ in range(reps):
    rc = w._lib.bmpc_ntt_dev(w.ctx, coeffs.data_ptr(), log_m, 0, st.cuda_stream)
    assert rc == 0
torch.cuda.synchronize()
print("ok", log_m)
