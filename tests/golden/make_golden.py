"""Generates tests/golden/vectors.json from the big-integer oracle (oracle/*.py, which is
pinned to the reference's own dummy-engine known-answer vectors).  The reference is Rust-only and
cannot be imported or built here, so these BLS12-381 vectors are oracle outputs on seeded inputs;
they freeze today's oracle behaviour and give the CUDA path fixed byte strings to reproduce.

    python tests/golden/make_golden.py
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import curves, domain, fields  # noqa: E402
from oracle import groth16 as og  # noqa: E402
from oracle import multiexp as ome  # noqa: E402

Q = fields.Fr.p
rng = random.Random(20261018)
out = {"ntt": [], "msm": [], "proof": []}

for logn in (0, 3, 6, 9):
    coeffs = [rng.randrange(Q) for _ in range(1 << logn)]
    rec = {"logn": logn, "coeffs": [hex(c) for c in coeffs]}
    for name in ("fft", "ifft", "coset_fft", "icoset_fft"):
        d = domain.EvaluationDomain(fields.Fr, coeffs)
        getattr(d, name)()
        rec[name] = [hex(c) for c in d.coeffs]
    out["ntt"].append(rec)

for gname, G, n, start in (("G1", curves.G1, 24, 2), ("G2", curves.G2, 10, 1)):
    pts = [G.mul(G.gen, rng.randrange(1, Q)) for _ in range(n + start + 1)]
    scalars = [rng.choice([0, 1, rng.randrange(Q), rng.randrange(1 << 40)]) for _ in range(n)]
    bits = [rng.random() < 0.7 for _ in range(n)]
    dens = ome.DensityTracker()
    dens.bv = bits
    full = ome.multiexp(G, pts, 0, ome.FullDensity(), scalars)
    sparse = ome.multiexp(G, pts, start, dens, scalars)
    out["msm"].append({"group": gname, "bases": G.to_uncompressed(pts[0]).hex()[:0] + "".join(G.to_uncompressed(p).hex() for p in pts),
                       "scalars": [hex(s) for s in scalars], "bits": [int(b) for b in bits], "start": start,
                       "full": G.to_uncompressed(full).hex(), "sparse": G.to_uncompressed(sparse).hex()})

E = og.BLS12
params = og.generate_random_parameters(E, og.xor_demo(None, None))      # alpha=6 beta=24 gamma=6 delta=24 tau=2
G1, G2 = curves.G1, curves.G2
crs = {"h": "".join(G1.to_uncompressed(p).hex() for p in params.h),
       "l": "".join(G1.to_uncompressed(p).hex() for p in params.l),
       "a": "".join(G1.to_uncompressed(p).hex() for p in params.a),
       "b_g1": "".join(G1.to_uncompressed(p).hex() for p in params.b_g1),
       "b_g2": "".join(G2.to_uncompressed(p).hex() for p in params.b_g2),
       "alpha_g1": G1.to_uncompressed(params.vk.alpha_g1).hex(), "beta_g1": G1.to_uncompressed(params.vk.beta_g1).hex(),
       "beta_g2": G2.to_uncompressed(params.vk.beta_g2).hex(), "delta_g1": G1.to_uncompressed(params.vk.delta_g1).hex(),
       "delta_g2": G2.to_uncompressed(params.vk.delta_g2).hex()}
for a, b in ((False, False), (True, False)):
    pr = og.synthesize_for_proving(E, og.xor_demo(a, b))
    proof = og.create_proof_from_assignment(E, pr, params, 27134, 17146)
    assert proof.to_bytes(E) == og.expected_proof(E, params, pr, 27134, 17146).to_bytes(E)
    out["proof"].append({"circuit": "XorDemo", "a": a, "b": b, "r": 27134, "s": 17146,
                         "eval_a": [hex(x) for x in pr.a], "eval_b": [hex(x) for x in pr.b], "eval_c": [hex(x) for x in pr.c],
                         "inputs": [hex(x) for x in pr.input_assignment], "aux": [hex(x) for x in pr.aux_assignment],
                         "a_aux_density": [int(x) for x in pr.a_aux_density.bv],
                         "b_input_density": [int(x) for x in pr.b_input_density.bv],
                         "b_aux_density": [int(x) for x in pr.b_aux_density.bv],
                         "proof": proof.to_bytes(E).hex()})
out["crs_xordemo"] = crs

# list_mul_matrix (mpc.rs:416-457): own generator so that the vectors above keep their values
from oracle import mpc as ompc  # noqa: E402
rng2 = random.Random(4160457)
out["list_mul_matrix"] = []
for gname, G, n in (("G1", curves.G1, 9), ("G2", curves.G2, 5)):
    lst = [G.mul(G.gen, rng2.randrange(1, Q)) for _ in range(n)]
    lst[1] = G.identity()
    matrix = []
    for i in range(n - 1):
        k = 0 if i == n - 3 else rng2.randrange(1, 4)         # an empty row: the reference stops there
        matrix.append([(rng2.choice([0, 1, Q - 1, rng2.randrange(Q), rng2.randrange(1 << 40)]), rng2.randrange(n))
                       for _ in range(k)])
    res = ompc.list_mul_matrix(G, lst, matrix)
    out["list_mul_matrix"].append({"group": gname, "list": "".join(G.to_uncompressed(p).hex() for p in lst),
                                   "matrix": [[[hex(cf), idx] for cf, idx in row] for row in matrix],
                                   "result": "".join(G.to_uncompressed(p).hex() for p in res)})
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "vectors.json"), "w"), indent=0)
print("wrote vectors.json")
