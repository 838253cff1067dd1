"""Synthetic satisfiable R1CS in CSR form: 16 inputs, one new aux variable per
constraint defined as (x_p + x_q) * x_r over earlier variables.  Times the device-side constraint
evaluation (bmpc_r1cs_eval) and key generation (bmpc_generate_parameters, N2) and checks
the evaluations exactly against the values the generator computed while building the witness."""
from __future__ import annotations

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import bellman_mpc_b200 as bm  # noqa: E402

Q = bm.FR_MODULUS


def build(log_n, seed=9):
    n = 1 << log_n
    ni = 16
    rs = np.random.RandomState(seed)
    w = [1] + [int(x) for x in rs.randint(1, 1 << 62, size=ni - 1)]
    p = np.empty(n, dtype=np.int64); q = np.empty(n, dtype=np.int64); r = np.empty(n, dtype=np.int64)
    a_ev, b_ev = [0] * n, [0] * n
    u = rs.random_sample((n, 3))
    pick = (u * (ni + np.arange(n))[:, None]).astype(np.int64)          # uniform over the earlier variables
    p[:], q[:], r[:] = pick[:, 0], pick[:, 1], pick[:, 2]
    pl, ql, rl = p.tolist(), q.tolist(), r.tolist()
    for j in range(n):
        pj, qj, rj = pl[j], ql[j], rl[j]
        a_ev[j] = (w[pj] + w[qj]) % Q
        b_ev[j] = w[rj]
        w.append(a_ev[j] * b_ev[j] % Q)
    one = bm.fr_to_mont([1])[0]
    ones = lambda k: np.tile(one, (k, 1))
    A = bm.CsrMatrix(np.arange(0, 2 * n + 1, 2), np.stack([p, q], axis=1).reshape(-1), ones(2 * n))
    B = bm.CsrMatrix(np.arange(0, n + 1), r, ones(n))
    C = bm.CsrMatrix(np.arange(0, n + 1), np.arange(ni, ni + n), ones(n))
    return dict(n=n, ni=ni, na=n, w=w, A=A, B=B, C=C, a=a_ev, b=b_ev, cols=(np.stack([p, q], axis=1).reshape(-1), r))


def transpose(csr, nv):
    """CSR of the transpose (index bookkeeping with numpy; coefficients are all ONE here)"""
    rows = np.repeat(np.arange(csr.num_rows, dtype=np.uint32), np.diff(csr.row_ptr).astype(np.int64))
    order = np.argsort(csr.col, kind="stable")
    counts = np.bincount(csr.col, minlength=nv)
    row_ptr = np.concatenate([[0], np.cumsum(counts)])
    return bm.CsrMatrix(row_ptr, rows[order], csr.coeff[order])


def run(worker, log_n=20, with_keygen=True):
    t0 = time.perf_counter()
    sysd = build(log_n)
    t_build = time.perf_counter() - t0
    n, ni, na = sysd["n"], sysd["ni"], sysd["na"]
    inputs = bm.fr_to_mont(sysd["w"][:ni])
    t0 = time.perf_counter()
    aux = bm.fr_to_mont(sysd["w"][ni:])
    t_mont = time.perf_counter() - t0
    bm.r1cs_eval(worker, sysd["A"], sysd["B"], sysd["C"], inputs, aux)            # warm-up
    t0 = time.perf_counter()
    asg = bm.r1cs_eval(worker, sysd["A"], sysd["B"], sysd["C"], inputs, aux)
    t_eval = time.perf_counter() - t0
    # exactness: a, b as computed by the generator; c = the aux value itself; densities from the columns
    idx = np.random.RandomState(1).randint(0, n, size=2000)
    ok = bm.fr_from_mont(asg.a[idx]) == [sysd["a"][i] for i in idx] and bm.fr_from_mont(asg.b[idx]) == [sysd["b"][i] for i in idx]
    ok = ok and np.array_equal(asg.c[:n], aux) and bm.fr_from_mont(asg.a[n:n + ni]) == sysd["w"][:ni]
    pq, r = sysd["cols"]
    exp_a_aux = np.zeros(na, dtype=bool); exp_a_aux[pq[pq >= ni] - ni] = True
    exp_b_aux = np.zeros(na, dtype=bool); exp_b_aux[r[r >= ni] - ni] = True
    exp_b_in = np.zeros(ni, dtype=bool); exp_b_in[r[r < ni]] = True
    ok = ok and np.array_equal(asg.a_aux_density.bv, exp_a_aux) and np.array_equal(asg.b_aux_density.bv, exp_b_aux) \
        and np.array_equal(asg.b_input_density.bv, exp_b_in)
    nnz = 4 * n
    res = {"constraints": n, "nnz": nnz, "r1cs_eval_s": t_eval, "r1cs_eval_mconstraints_per_s": n / t_eval / 1e6,
           "matches_generator": bool(ok), "host_build_s": round(t_build, 2), "host_to_mont_s": round(t_mont, 2),
           "note": "host CSR + assignments -> H2D -> 3 SpMV + densities -> D2H a, b, c (prover.rs:19-53,100-138,202-204)"}
    if with_keygen:
        nv = ni + na
        t0 = time.perf_counter()
        T = [transpose(M, nv) for M in (sysd["A"], sysd["B"], sysd["C"])]
        t_tr = time.perf_counter() - t0
        from bench_prove import G1_GEN, G2_GEN
        t0 = time.perf_counter()
        gp = bm.generate_parameters(worker, *T, ni, na, n, G1_GEN, G2_GEN, 6, 24, 6, 24, 2)
        t_kg = time.perf_counter() - t0
        res.update({"keygen_s": t_kg, "keygen_transpose_host_s": round(t_tr, 2),
                    "keygen_sizes": {"h": len(gp.h), "l": len(gp.l), "a": len(gp.a), "b_g1": len(gp.b_g1), "b_g2": len(gp.b_g2)}})
        # the generated CRS proves the generated assignment; the proof is deterministic
        t0 = time.perf_counter()
        p1 = bm.create_random_proof(asg, gp)
        res["prove_with_generated_crs_s"] = time.perf_counter() - t0
        res["proof_deterministic"] = bool(p1 == bm.create_random_proof(asg, gp))
        gp.free()
    return res


if __name__ == "__main__":
    import json
    w = bm.Worker(0)
    print(json.dumps(run(w, int(sys.argv[1]) if len(sys.argv) > 1 else 20)))
