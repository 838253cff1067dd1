"""The oracle's BLS12-381 pairing and the reference's verification equation (verifier.rs:10-62) on
it.  CPU only.  The pairing is pinned by the properties every use in the reference relies on
(bilinearity, order r, non-degeneracy: oracle/pairing.py self_check) and is at the same time an
independent check of oracle/curves.py: a wrong group law in G1 or G2 breaks e(aP, bQ) = e(P, Q)^(ab)."""
import random

import pytest

from oracle import curves, fields, pairing
from oracle import groth16 as og

Q = fields.Fr.p


def test_pairing_self_check():
    assert pairing.self_check()


def test_bilinearity_random_scalars():
    rng = random.Random(7)
    a, b = rng.randrange(Q), rng.randrange(Q)
    G1, G2 = curves.G1, curves.G2
    e = pairing.pairing(G1.gen, G2.gen)
    assert pairing.pairing(G1.mul(G1.gen, a), G2.mul(G2.gen, b)) == pairing.f12_pow(e, a * b % Q)
    # additivity in the first argument through the product of Miller loops (multi_miller_loop)
    P1, P2 = G1.mul(G1.gen, a), G1.mul(G1.gen, b)
    lhs = pairing.final_exponentiation(pairing.multi_miller_loop([(P1, G2.gen), (P2, G2.gen)]))
    assert lhs == pairing.pairing(G1.add(P1, P2), G2.gen)


def test_point_compression_round_trip_and_rejects():
    G1, G2 = curves.G1, curves.G2
    assert curves.self_check()
    good = bytearray(G1.to_compressed(G1.mul(G1.gen, 12345)))
    assert G1.from_compressed(bytes(good)) == G1.mul(G1.gen, 12345)
    bad = bytearray(good)
    bad[0] &= 0x7F                                   # compression flag missing
    assert G1.from_compressed(bytes(bad)) is None
    # x with no point above it / a point outside the r-torsion: find one by stepping x
    x = 5
    while True:
        enc = bytearray(x.to_bytes(48, "big"))
        enc[0] |= 0x80
        pt = G1.from_compressed(bytes(enc))
        rhs = (x ** 3 + 4) % fields.FP_MODULUS
        if curves.fp_sqrt(rhs) is not None:
            assert pt is None                        # on the curve, not in the subgroup (cofactor ~2^126)
            break
        assert pt is None
        x += 1
    assert G2.from_compressed(G2.to_compressed(G2.mul(G2.gen, 99))) == G2.mul(G2.gen, 99)


def test_verify_xor_demo_bls12():
    """groth16/tests/mod.rs XorDemo through generate -> create_proof -> Proof::write/read ->
    prepare_verifying_key -> verify_proof on the real curve"""
    E = og.BLS12
    params = og.generate_random_parameters(E, og.xor_demo(None, None))
    pvk = og.prepare_verifying_key(E, params.vk)
    prover = og.synthesize_for_proving(E, og.xor_demo(True, False))
    proof = og.create_proof_from_assignment(E, prover, params, 27134, 17146)
    back = og.Proof.read(E, proof.to_bytes(E))
    assert (back.a, back.b, back.c) == (proof.a, proof.b, proof.c)
    assert og.verify_proof_prepared(E, pvk, back, [1])
    assert not og.verify_proof_prepared(E, pvk, back, [0])             # wrong public input
    with pytest.raises(og.InvalidVerifyingKey):                          # verifier.rs:28-30
        og.verify_proof_prepared(E, pvk, back, [])
    forged = og.Proof(proof.a, proof.b, curves.G1.add(proof.c, curves.G1.gen))
    assert not og.verify_proof_prepared(E, pvk, forged, [1])
    # the known-trapdoor expectation used by the GPU tests is a proof the verifier accepts
    exp = og.expected_proof(E, params, prover, 27134, 17146)
    assert og.verify_proof_prepared(E, pvk, exp, [1])


def test_proof_read_errors():
    E = og.BLS12
    G1, G2 = curves.G1, curves.G2
    a, b = G1.to_compressed(G1.gen), G2.to_compressed(G2.gen)
    with pytest.raises(ValueError, match="UnexpectedEof"):
        og.Proof.read(E, (a + b + a)[:-1])
    with pytest.raises(ValueError, match="point at infinity"):
        og.Proof.read(E, G1.to_compressed(None) + b + a)
    with pytest.raises(ValueError, match="invalid G2"):
        og.Proof.read(E, a + bytes([b[0] & 0x7F]) + b[1:] + a)
