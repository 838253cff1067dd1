"""Pins the oracle to the reference's own known-answer vectors: the DummyEngine
(F_64513) values in src/groth16/tests/mod.rs and dummy_engine.rs, the Fp constants in
src/gt_bytes.rs, plus the property tests the reference runs on BLS12-381."""
import random

from oracle import curves, domain, fields
from oracle import groth16 as g
from oracle import multiexp as me


def test_constants_self_check():
    assert fields.self_check() and curves.self_check()


def test_xordemo_golden_vectors():
    """groth16/tests/mod.rs:299-589 (test_xordemo)"""
    E = g.DUMMY
    alpha, beta, gamma, delta, tau = 48577, 22580, 53332, 5481, 3673
    params = g.generate_parameters(E, g.xor_demo(None, None), 1, 1, alpha, beta, gamma, delta, tau)
    p = 64513
    assert len(params.h) == 7                                   # :333
    root = pow(57751, 1 << 7, p)
    assert root == 20201                                        # :342
    t_at_tau = (pow(tau, 8, p) - 1) % p
    coeff = pow(delta, -1, p) * t_at_tau % p
    assert params.h == [pow(tau, i, p) * coeff % p for i in range(7)]   # :366-381
    assert (len(params.vk.ic), len(params.l), len(params.a), len(params.b_g1), len(params.b_g2)) == (2, 2, 4, 2, 2)  # :383-394
    u_i, v_i, w_i = [59158, 48317, 21767, 10402], [0, 0, 60619, 30791], [0, 23320, 41193, 41193]   # :424-435
    assert params.a == u_i
    assert params.b_g1 == [v for v in v_i if v] and params.b_g2 == [v for v in v_i if v]
    for i in range(4):                                          # :457-478
        t = (beta * u_i[i] + alpha * v_i[i] + w_i[i]) % p
        if i < 2:
            assert params.vk.ic[i] == t * pow(gamma, -1, p) % p
        else:
            assert params.l[i - 2] == t * pow(delta, -1, p) % p
    assert (params.vk.alpha_g1, params.vk.beta_g1, params.vk.beta_g2, params.vk.gamma_g2,
            params.vk.delta_g1, params.vk.delta_g2) == (alpha, beta, beta, gamma, delta, delta)
    r, s = 27134, 17146
    # upstream instance a=1, b=0: quotient coefficients pinned at :574
    prover = g.synthesize_for_proving(E, g.xor_demo(True, False))
    assert g.h_coefficients(E.Fr, prover.a, prover.b, prover.c) == [5040, 11763, 10755, 63633, 128, 9747, 8739]
    proof = g.create_proof_from_assignment(E, prover, params, r, s)
    assert proof.a == (delta * r + alpha + u_i[0] + u_i[1] + u_i[2]) % p        # :497-508
    assert proof.b == (delta * s + beta + v_i[0] + v_i[1] + v_i[2]) % p         # :517-528
    exp_c = (proof.a * s + proof.b * r - delta * r * s + params.l[0]
             + sum(h * c for h, c in zip(params.h, [5040, 11763, 10755, 63633, 128, 9747, 8739]))) % p
    assert proof.c == exp_c                                                      # :548-585
    assert g.verify_proof(E, params.vk, proof, [1])
    assert not g.verify_proof(E, params.vk, proof, [0])
    # the fork's instance a=0, b=0 (:494-497, :587)
    prover0 = g.synthesize_for_proving(E, g.xor_demo(False, False))
    proof0 = g.create_proof_from_assignment(E, prover0, params, r, s)
    assert g.verify_proof(E, params.vk, proof0, [0])
    ex = g.expected_proof(E, params, prover0, r, s)
    assert (ex.a, ex.b, ex.c) == (proof0.a, proof0.b, proof0.c)


def test_multiexp_matches_naive_dummy_and_bls():
    """multiexp.rs:283-327 (test_with_bls12) at oracle-friendly sizes, plus the dummy group"""
    rng = random.Random(1)
    for G, n in ((curves.Dummy, 500), (curves.G1, 40), (curves.G2, 33)):
        F = G.scalar_field
        bases = [G.mul(G.gen, rng.randrange(1, F.p)) for _ in range(n)]
        exps = [rng.randrange(F.p) for _ in range(n)]
        assert G.eq(me.multiexp(G, bases, 0, me.FullDensity(), exps), me.naive(G, bases, exps))


def test_multiexp_density_offset_errors():
    G, F = curves.Dummy, fields.DummyFr
    rng = random.Random(2)
    n = 64
    bases = [rng.randrange(1, F.p) for _ in range(n + 5)]
    exps = [rng.choice([0, 1, rng.randrange(F.p)]) for _ in range(n)]
    d = me.DensityTracker()
    d.bv = [rng.random() < 0.5 for _ in range(n)]
    k, acc = 3, 0
    for e, bit in zip(exps, d.bv):
        if bit:
            acc = (acc + bases[k] * e) % F.p
            k += 1
    assert me.multiexp(G, bases, 3, d, exps) == acc
    assert d.to_words()[0] == sum(1 << i for i in range(64) if d.bv[i])
    import pytest
    with pytest.raises(me.UnexpectedEof):
        me.multiexp(G, bases[:10], 0, me.FullDensity(), exps)
    b2 = list(bases)
    b2[0] = 0
    e2 = list(exps)
    e2[0] = 5
    with pytest.raises(me.UnexpectedIdentity):
        me.multiexp(G, b2, 0, me.FullDensity(), e2)
    e2[0] = 0
    me.multiexp(G, b2, 0, me.FullDensity(), e2)
    assert me.multiexp(G, bases, 999, me.FullDensity(), []) == 0
    assert [me.window_size(n) for n in (2, 31, 32, 645, 1023, 1 << 14, 1 << 16, 1 << 20, 1 << 22, 1 << 24, 1 << 26)] \
        == [3, 3, 4, 7, 7, 10, 12, 14, 16, 17, 19]


def test_domain_properties():
    """domain.rs:376-498: polynomial_arith, fft_composition, parallel_fft_consistency"""
    rng = random.Random(3)
    for F in (fields.DummyFr, fields.Fr):
        maxlog = 9 if F is fields.DummyFr else 6
        for logn in range(0, maxlog):
            n = 1 << logn
            c = [rng.randrange(F.p) for _ in range(n)]
            d = domain.EvaluationDomain(F, c)
            assert d.omega == pow(F.root_of_unity, 1 << (F.S - logn), F.p)
            d.fft()
            assert d.coeffs == domain.dft_by_definition(F, c, d.omega)
            for f, gname in (("ifft", "fft"), ("fft", "ifft"), ("icoset_fft", "coset_fft"), ("coset_fft", "icoset_fft")):
                e = domain.EvaluationDomain(F, c)
                getattr(e, f)()
                getattr(e, gname)()
                assert e.coeffs == c
            for log_cpus in range(0, min(3, logn + 1)):
                e = domain.EvaluationDomain(F, c, log_cpus=log_cpus)
                e.fft()
                assert e.coeffs == d.coeffs
    F = fields.Fr
    a = [rng.randrange(F.p) for _ in range(13)]
    b = [rng.randrange(F.p) for _ in range(20)]
    naive = [0] * 33
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            naive[i + j] = (naive[i + j] + x * y) % F.p
    da, db = domain.EvaluationDomain(F, a + [0] * 20), domain.EvaluationDomain(F, b + [0] * 13)
    da.fft(); db.fft(); da.mul_assign(db); da.ifft()
    assert da.coeffs[:33] == naive and not any(da.coeffs[33:])
    import pytest
    with pytest.raises(me.PolynomialDegreeTooLarge):
        domain.EvaluationDomain(fields.DummyFr, [0] * 1025)    # exp >= S = 10


def test_bls_proof_matches_known_trapdoor_and_sizes():
    """groth16/mod.rs:489-570: proof = 192 B; create_proof == known-trapdoor algebra"""
    E = g.BLS12
    params = g.generate_random_parameters(E, g.xor_demo(None, None))
    prover = g.synthesize_for_proving(E, g.xor_demo(True, True))
    proof = g.create_random_proof(E, g.xor_demo(True, True), params)
    assert len(proof.to_bytes(E)) == 192
    assert proof.to_bytes(E) == g.expected_proof(E, params, prover, 27134, 17146).to_bytes(E)
