"""GPU parity: create_proof (src/groth16/prover.rs:176-350) on the reference's own demo
circuits, checked byte-for-byte against the oracle restatement and against the
known-trapdoor expectation."""
import random

import numpy as np
import pytest

import bellman_mpc_b200 as bm
from oracle import curves, fields
from oracle import groth16 as og

pytestmark = pytest.mark.gpu
F = fields.Fr
Q = F.p


def upload_params(worker, params):
    G1, G2 = curves.G1, curves.G2
    up1 = lambda v: bm.Bases.from_uncompressed(worker, bm.G1, b"".join(G1.to_uncompressed(p) for p in v), len(v))
    up2 = lambda v: bm.Bases.from_uncompressed(worker, bm.G2, b"".join(G2.to_uncompressed(p) for p in v), len(v))
    vk = params.vk
    return bm.Parameters(worker, up1(params.h), up1(params.l), up1(params.a), up1(params.b_g1), up2(params.b_g2),
                         G1.to_uncompressed(vk.alpha_g1), G1.to_uncompressed(vk.beta_g1),
                         G2.to_uncompressed(vk.beta_g2), G1.to_uncompressed(vk.delta_g1),
                         G2.to_uncompressed(vk.delta_g2))


def to_gpu_assignment(prover):
    dens = lambda d: bm.DensityTracker.from_bits(d.bv)
    return bm.ProvingAssignment(bm.fr_to_mont(prover.a), bm.fr_to_mont(prover.b), bm.fr_to_mont(prover.c),
                                bm.fr_to_mont(prover.input_assignment), bm.fr_to_mont(prover.aux_assignment),
                                dens(prover.a_aux_density), dens(prover.b_input_density),
                                dens(prover.b_aux_density))


def test_xor_demo(worker):
    """groth16/tests/mod.rs XorDemo on BLS12-381 with the fork's fixed toxic waste and r, s"""
    E = og.BLS12
    params = og.generate_random_parameters(E, og.xor_demo(None, None))
    gp = upload_params(worker, params)
    for a, b in [(False, False), (True, False), (True, True)]:
        prover = og.synthesize_for_proving(E, og.xor_demo(a, b))
        proof = bm.create_random_proof(to_gpu_assignment(prover), gp)
        assert len(proof) == 192
        assert proof == og.create_proof_from_assignment(E, prover, params, 27134, 17146).to_bytes(E)
        assert proof == og.expected_proof(E, params, prover, 27134, 17146).to_bytes(E)
        # the reference verifier (verifier.rs:23-62) accepts the bytes the GPU produced
        got = og.Proof.read(E, proof)
        assert og.verify_proof(E, params.vk, got, [int(a ^ b)])
        assert not og.verify_proof(E, params.vk, got, [int(not (a ^ b))])


def test_full_size_r_s(worker):
    """create_proof(circuit, params, r, s) with 255-bit r, s (prover.rs:176-181 takes them as arguments;
    only create_random_proof fixes them): the seven scalar multiplications of the tail walk all 255 bits.
    Bytes == the oracle prover's, and the verifier accepts."""
    E = og.BLS12
    rng = random.Random(99)
    params = og.generate_random_parameters(E, og.xor_demo(None, None))
    gp = upload_params(worker, params)
    prover = og.synthesize_for_proving(E, og.xor_demo(True, True))
    for r, s in ((rng.randrange(Q), rng.randrange(Q)), (Q - 1, 1), (0, 0)):
        proof = bm.create_proof(to_gpu_assignment(prover), gp, bm.fr_to_mont([r])[0], bm.fr_to_mont([s])[0])
        assert proof == og.create_proof_from_assignment(E, prover, params, r, s).to_bytes(E), (r, s)
        assert og.verify_proof(E, params.vk, og.Proof.read(E, proof), [0])


def test_mimc(worker):
    """config #1: the crate's MiMC circuit (src/mimc_mod.rs; 646 constraints -> domain 1024,
    MSM sizes 1023 / 645 / 2), proof bytes == known-trapdoor proof"""
    E = og.BLS12
    rng = random.Random(2024)
    constants = [rng.randrange(Q) for _ in range(og.MIMC_ROUNDS)]
    xl, xr = rng.randrange(Q), rng.randrange(Q)
    params = og.generate_random_parameters(E, og.mimc_demo(F, None, None, constants))
    assert len(params.h) == 1023 and len(params.l) == 645
    gp = upload_params(worker, params)
    prover = og.synthesize_for_proving(E, og.mimc_demo(F, xl, xr, constants))
    assert prover.input_assignment[1] == og.mimc(F, xl, xr, constants)
    assert prover.a_aux_density.get_total_density() == 644
    assert prover.b_aux_density.get_total_density() == 322
    proof = bm.create_random_proof(to_gpu_assignment(prover), gp)
    assert proof == og.expected_proof(E, params, prover, 27134, 17146).to_bytes(E)
    # config #1 is "prove + verify": Proof::read, prepare_verifying_key, verify_proof (tests/mimc.rs)
    pvk = og.prepare_verifying_key(E, params.vk)
    got = og.Proof.read(E, proof)
    image = og.mimc(F, xl, xr, constants)
    assert og.verify_proof_prepared(E, pvk, got, [image])
    assert not og.verify_proof_prepared(E, pvk, got, [(image + 1) % Q])


def test_synthetic_2p14_proof_accepted_by_the_verifier(worker):
    """A satisfiable random R1CS with 2^14 constraints (16 inputs, one aux variable per constraint):
    key generation, constraint evaluation and create_proof all on the GPU; the 192 proof bytes go
    through Proof::read and the reference's verification equation with the oracle's pairing.  Needs no
    known-dlog expectation: acceptance is an independent statement about H, the four query
    multiexps and the tail."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench"))
    import r1cs_bench
    from oracle import params_io as pio
    E = og.BLS12
    sysd = r1cs_bench.build(14, seed=11)
    n, ni, na = sysd["n"], sysd["ni"], sysd["na"]
    asg = bm.r1cs_eval(worker, sysd["A"], sysd["B"], sysd["C"], bm.fr_to_mont(sysd["w"][:ni]),
                       bm.fr_to_mont(sysd["w"][ni:]))
    T = [r1cs_bench.transpose(M, ni + na) for M in (sysd["A"], sysd["B"], sysd["C"])]
    G1, G2 = curves.G1, curves.G2
    gp = bm.generate_parameters(worker, *T, ni, na, n, G1.to_uncompressed(G1.gen), G2.to_uncompressed(G2.gen),
                                6, 24, 6, 24, 2)
    proof = bm.create_random_proof(asg, gp)
    vk = pio.read_vk(pio._Reader(gp.write()))
    assert len(vk.ic) == ni
    pvk = og.prepare_verifying_key(E, vk)
    got = og.Proof.read(E, proof)
    public = sysd["w"][1:ni]
    assert og.verify_proof_prepared(E, pvk, got, public)
    wrong = list(public)
    wrong[3] = (wrong[3] + 1) % Q
    assert not og.verify_proof_prepared(E, pvk, got, wrong)
    gp.free()


def test_delta_identity_rejected(worker):
    """prover.rs:309-313 subversion check"""
    E = og.BLS12
    params = og.generate_random_parameters(E, og.xor_demo(None, None))
    gp = upload_params(worker, params)
    gp.delta_g1 = curves.G1.to_uncompressed(None)
    prover = og.synthesize_for_proving(E, og.xor_demo(True, False))
    with pytest.raises(bm.UnexpectedIdentity):
        bm.create_random_proof(to_gpu_assignment(prover), gp)


@pytest.mark.parametrize("log_m,profile", [(5, "uniform"), (12, "uniform"), (14, "boolean")])
def test_synthetic_prove_known_dlog_and_cpu(worker, log_m, profile):
    """config #4 at test sizes: synthetic R1CS state with random densities and a known-dlog CRS.
    GPU proof == known-dlog expectation == C restatement of prover.rs:206-350."""
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench_prove import Workload
    from oracle import cref
    wl = Workload(worker, log_m, seed=40 + log_m, profile=profile)
    proof = wl.prove()
    assert proof == wl.expected_proof()
    st, cpu_proof, _ = wl.cpu_reference_proof(threads=4)
    assert st == 0 and cpu_proof == proof
    wl.free()


@pytest.mark.parametrize("world", [1, 2, 3])
def test_sharded_create_proof_emulated(worker, world):
    """create_proof split over `world` ranks (each holding only its slices of the query
    vectors) gives the same 192 bytes as the single-GPU call.  The ranks run one after the other on
    this GPU through the same C-ABI entry points the multi-process path uses
    (bmpc_create_proof_partials / bmpc_create_proof_finish); the collective is a list append."""
    import bench_prove
    from bellman_mpc_b200 import dist as bdist
    log_m = 10
    full = bench_prove.Workload(worker, log_m)
    expect = full.prove()
    assert expect == full.expected_proof()
    parts, stats, keep = [], [], []
    for rank in range(world):
        wl = bench_prove.Workload(worker, log_m, world=world, rank=rank)
        pb, st = bdist.proof_partials(worker, wl.params, wl.assignment, wl.plan)
        assert pb is not None, st
        parts.append(pb)
        stats.append(st)
        keep.append(wl)
    rc, proof = bdist.proof_finish(worker, keep[0].params, parts, stats, full.r, full.s)
    assert rc == 0
    assert proof == expect
    for wl in keep:
        wl.free()
    full.free()


@pytest.mark.parametrize("world", [2, 3, 5])
def test_sharded_create_proof_shared_h_emulated(worker, world):
    """The H polynomial computed once by the ranks together (dist.HSplit: one of a, b, c per rank 0..2,
    combined on rank 0, every rank gets its slice of the scalars and runs its share of the H multiexp) and
    the other seven multiexps per rank without H: same 192 bytes as the single-GPU call.  Ranks run one
    after the other on this GPU; sends and receives are tensor hand-overs."""
    import torch
    import bench_prove
    from bellman_mpc_b200 import dist as bdist
    log_m = 10
    full = bench_prove.Workload(worker, log_m)
    expect = full.prove()
    assert expect == full.expected_proof()
    hw = bm.Worker(0)
    wls = [bench_prove.Workload(worker, log_m, world=world, rank=r) for r in range(world)]
    m = wls[0].plan.m
    # coset evaluations by their owners, then rank 0 combines
    ev = [None] * 3
    for k, name in enumerate(("a", "b", "c")):
        o = bdist.h_owner(k, world)
        hs = bdist.HSplit(hw, wls[o].plan)
        ev[k] = hs.coset_evals(getattr(wls[o].assignment, name), torch.empty((m, 4), dtype=torch.int64, device="cuda"))
    h_all = bdist.HSplit(hw, wls[0].plan).combine(ev[0], ev[1], ev[2])
    torch.cuda.synchronize()
    # == the host-buffer H pipeline
    want_h = bm.h_coefficients(worker, full.a, full.b, full.c)
    assert np.array_equal(h_all[:m - 1].cpu().numpy().view(np.uint64), want_h)
    parts, flags = [], []
    for r, wl in enumerate(wls):
        pb, fl = bdist.proof_partials(worker, wl.params, wl.assignment, wl.plan, skip_h=True)
        assert pb is not None, fl
        hp = torch.zeros(192, dtype=torch.uint8, device="cuda")
        rc, fh = bdist.HSplit(hw, wl.plan).h_partial(wl.params.h, h_all[wl.plan.h_lo:wl.plan.h_hi].contiguous(), hp)
        assert rc == 0
        pb, fl = bdist.splice_h(pb, fl, bytes(hp.cpu().numpy().tobytes()), fh)
        parts.append(pb)
        flags.append(fl)
    rc, proof = bdist.proof_finish(worker, wls[0].params, parts, flags, full.r, full.s)
    assert rc == 0 and proof == expect
    for wl in wls:
        wl.free()
    full.free()
    hw.close()
