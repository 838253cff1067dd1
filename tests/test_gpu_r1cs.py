"""GPU parity for the constraint-system side: bmpc_r1cs_eval against the oracle's
ProvingAssignment (prover.rs:19-156), bmpc_generate_parameters against the oracle's upstream-semantics
generate_parameters (generator.rs:241-634) byte-for-byte on the serialised Parameters, and the
chain keygen -> evaluation -> create_proof against the known-trapdoor proof."""
import random

import numpy as np
import pytest

import bellman_mpc_b200 as bm
from oracle import curves, fields
from oracle import groth16 as og
from oracle import params_io as pio

pytestmark = pytest.mark.gpu
Q = fields.Fr.p


class Recorder:
    """records a circuit's constraints as rows of (variable, coeff) -- the R1CS a host passes in CSR form"""

    def __init__(self):
        self.inputs, self.aux, self.rows = [], [], []

    def alloc(self, f):
        self.aux.append(f())
        return ("aux", len(self.aux) - 1)

    def alloc_input(self, f):
        self.inputs.append(f())
        return ("input", len(self.inputs) - 1)

    def enforce(self, a, b, c):
        self.rows.append((a, b, c))


def record(circuit):
    rec = Recorder()
    rec.alloc_input(lambda: 1)
    circuit(rec)
    ni = len(rec.inputs)
    col = lambda v: v[1] if v[0] == "input" else ni + v[1]
    mats = []
    for k in range(3):
        mats.append([[(col(v), c) for v, c in row[k]] for row in rec.rows])
    return rec, mats


def transpose(rows, nv):
    cols = [[] for _ in range(nv)]
    for r, row in enumerate(rows):
        for c, v in row:
            cols[c].append((r, v))
    return cols


CIRCUITS = {
    "xor": lambda: og.xor_demo(True, False),
    "silly": lambda: pio.my_silly_circuit(3, 5),
    "mimc": None,
}


def _mimc():
    rng = random.Random(2024)
    constants = [rng.randrange(Q) for _ in range(og.MIMC_ROUNDS)]
    return og.mimc_demo(fields.Fr, rng.randrange(Q), rng.randrange(Q), constants)


@pytest.mark.parametrize("name", ["xor", "silly", "mimc"])
def test_eval_keygen_prove(worker, name):
    circuit = _mimc() if name == "mimc" else CIRCUITS[name]()
    E = og.BLS12
    rec, (A, B, C_) = record(circuit)
    ni, na, nc = len(rec.inputs), len(rec.aux), len(rec.rows)
    # --- N4: evaluation + densities
    prover = og.synthesize_for_proving(E, circuit)
    asg = bm.r1cs_eval(worker, bm.CsrMatrix.from_rows(A), bm.CsrMatrix.from_rows(B), bm.CsrMatrix.from_rows(C_),
                       bm.fr_to_mont(rec.inputs), bm.fr_to_mont(rec.aux))
    assert bm.fr_from_mont(asg.a) == prover.a and bm.fr_from_mont(asg.b) == prover.b and bm.fr_from_mont(asg.c) == prover.c
    assert list(asg.a_aux_density.bv) == prover.a_aux_density.bv
    assert list(asg.b_input_density.bv) == prover.b_input_density.bv
    assert list(asg.b_aux_density.bv) == prover.b_aux_density.bv
    # --- N2: key generation (the fork's fixed toxic waste, generator.rs:34-38)
    params = og.generate_random_parameters(E, circuit)
    nv = ni + na
    gp = bm.generate_parameters(worker, *(bm.CsrMatrix.from_rows(transpose(M, nv)) for M in (A, B, C_)), ni, na, nc,
                                curves.G1.to_uncompressed(curves.G1.gen), curves.G2.to_uncompressed(curves.G2.gen),
                                6, 24, 6, 24, 2)
    assert gp.write() == pio.write_parameters(params)
    # --- chain: GPU keygen -> GPU evaluation -> GPU proof == known-trapdoor proof
    proof = bm.create_random_proof(asg, gp)
    assert proof == og.expected_proof(E, params, prover, 27134, 17146).to_bytes(E)
    gp.free()


def test_keygen_errors(worker):
    rec, (A, B, C_) = record(og.xor_demo(True, True))
    ni, na, nc = len(rec.inputs), len(rec.aux), len(rec.rows)
    T = [bm.CsrMatrix.from_rows(transpose(M, ni + na)) for M in (A, B, C_)]
    g1, g2 = curves.G1.to_uncompressed(curves.G1.gen), curves.G2.to_uncompressed(curves.G2.gen)
    with pytest.raises(bm.UnexpectedIdentity):                       # delta = 0, generator.rs:338-345
        bm.generate_parameters(worker, *T, ni, na, nc, g1, g2, 6, 24, 6, 0, 2)
    # an aux variable that appears in no constraint -> UnconstrainedVariable (generator.rs:584-590)
    T2 = [bm.CsrMatrix.from_rows(transpose(M, ni + na) + [[]]) for M in (A, B, C_)]
    with pytest.raises(bm.InvalidData):
        bm.generate_parameters(worker, *T2, ni, na + 1, nc, g1, g2, 6, 24, 6, 24, 2)
