"""Pins the C restatement (oracle/c) to the Python oracle (which is pinned to the reference's
golden vectors): same algorithm, same inputs, byte-identical outputs."""
import random

import numpy as np

from oracle import cref, curves, domain, fields
from oracle import groth16 as og
from oracle import multiexp as ome

Q = fields.Fr.p
R = fields.Fr.R


def limbs(vals):
    out = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        for j in range(4):
            out[i, j] = (v >> (64 * j)) & (2**64 - 1)
    return out


def ints(arr):
    return [sum(int(x) << (64 * j) for j, x in enumerate(row)) for row in np.asarray(arr).reshape(-1, 4)]


def mont(vals):
    return limbs([v * R % Q for v in vals])


def unmont(arr):
    ri = pow(R, -1, Q)
    return [v * ri % Q for v in ints(arr)]


def words(bits):
    w = [0] * max(1, (len(bits) + 63) // 64)
    for i, b in enumerate(bits):
        if b:
            w[i // 64] |= 1 << (i % 64)
    return np.array(w, dtype=np.uint64)


def test_ntt_ops_match_python():
    rng = random.Random(1)
    for logn in (0, 1, 3, 6, 9):
        c = [rng.randrange(Q) for _ in range(1 << logn)]
        for op, name in enumerate(("fft", "ifft", "coset_fft", "icoset_fft")):
            o = domain.EvaluationDomain(fields.Fr, c)
            getattr(o, name)()
            for threads in (1, 4):
                assert unmont(cref.ntt(mont(c), op, threads)) == o.coeffs, (logn, name, threads)


def test_multiexp_matches_python_incl_errors():
    rng = random.Random(2)
    for G, grp, n in ((curves.G1, 1, 70), (curves.G2, 2, 40)):
        pts = [G.mul(G.gen, rng.randrange(1, Q)) for _ in range(n + 4)]
        pts[5] = None
        exps = [rng.choice([0, 1, rng.randrange(Q), rng.randrange(1 << 20)]) for _ in range(n)]
        exps[5] = 0
        bits = [rng.random() < 0.6 for _ in range(n)]
        cb = cref.CBases.from_uncompressed(grp, b"".join(G.to_uncompressed(p) for p in pts))
        for start, b in ((0, None), (2, bits)):
            dens = ome.FullDensity()
            if b is not None:
                dens = ome.DensityTracker()
                dens.bv = b
            e = list(exps)
            if b is not None:       # keep the identity base (index 5) under a zero scalar
                k = start
                for i, bit in enumerate(b):
                    if bit:
                        if k == 5:
                            e[i] = 0
                        k += 1
            exp = ome.multiexp(G, pts, start, dens, e)
            st, got = cref.multiexp(cb, start, limbs(e), None if b is None else words(b), threads=3)
            assert st == 0 and got == G.to_uncompressed(exp)
        assert cref.naive_multiexp(cb, limbs([3] * 5)) == G.to_uncompressed(ome.naive(G, pts[:5], [3] * 5))
        # errors
        e = list(exps); e[5] = 7
        assert cref.multiexp(cb, 0, limbs(e))[0] == 1
        assert cref.multiexp(cb, 10, limbs(exps))[0] == 2
        cb.free()
    assert [cref.load().orc_window_size(n) for n in (2, 32, 645, 1 << 20, 1 << 24)] == [3, 4, 7, 14, 17]


def test_h_and_proof_match_python():
    E = og.BLS12
    rng = random.Random(3)
    for n in (1, 5, 37):
        a, b, c = ([rng.randrange(Q) for _ in range(n)] for _ in range(3))
        assert ints(cref.h_coefficients(mont(a), mont(b), mont(c), threads=2)) == og.h_coefficients(fields.Fr, a, b, c)
    params = og.generate_random_parameters(E, og.xor_demo(None, None))
    G1, G2 = curves.G1, curves.G2
    mk = lambda G, grp, v: cref.CBases.from_uncompressed(grp, b"".join(G.to_uncompressed(p) for p in v))
    cp = cref.CParams()
    hs = [mk(G1, 1, params.h), mk(G1, 1, params.l), mk(G1, 1, params.a), mk(G1, 1, params.b_g1), mk(G2, 2, params.b_g2)]
    cp.h, cp.l, cp.a, cp.b_g1, cp.b_g2 = (h.handle for h in hs)
    import ctypes
    for name, G, pt in (("alpha_g1", G1, params.vk.alpha_g1), ("beta_g1", G1, params.vk.beta_g1),
                        ("beta_g2", G2, params.vk.beta_g2), ("delta_g1", G1, params.vk.delta_g1),
                        ("delta_g2", G2, params.vk.delta_g2)):
        raw = G.to_uncompressed(pt)
        ctypes.memmove(getattr(cp, name), raw, len(raw))
    for ab in ((False, False), (True, False)):
        pr = og.synthesize_for_proving(E, og.xor_demo(*ab))
        st, proof = cref.create_proof(cp, mont(pr.a), mont(pr.b), mont(pr.c), mont(pr.input_assignment),
                                      mont(pr.aux_assignment), words(pr.a_aux_density.bv),
                                      words(pr.b_input_density.bv), words(pr.b_aux_density.bv),
                                      mont([27134])[0], mont([17146])[0], threads=4)
        assert st == 0
        assert proof == og.expected_proof(E, params, pr, 27134, 17146).to_bytes(E)


def test_dot_and_generator_mul():
    rng = random.Random(4)
    k = [rng.randrange(Q) for _ in range(50)]
    s = [rng.randrange(Q) for _ in range(50)]
    d = ints(cref.fr_dot(limbs(k), limbs(s)))[0]
    assert d == sum(x * y for x, y in zip(k, s)) % Q
    assert cref.g1_generator_mul(limbs([d])[0]) == curves.G1.to_uncompressed(curves.G1.mul(curves.G1.gen, d))
