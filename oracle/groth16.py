"""TEST INFRASTRUCTURE ONLY -- restatement of the reference's Groth16 prover (and the
upstream-semantics key generator needed to feed it).

Follows /root/reference/bellman/src/groth16/:
  prover.rs:19-53    eval (LC evaluation + density tracking)
  prover.rs:55-156   ProvingAssignment
  prover.rs:158-173  create_random_proof (r = 27134, s = 17146, RNG ignored)
  prover.rs:176-350  create_proof
  mod.rs:224-247,438-477  Parameters / ParameterSource for &Parameters
  generator.rs:44-156     KeypairAssembly
  generator.rs:241-272,294-297,310-572,584-590,594-604,612-634
                          generate_parameters with the fork's MPC cross-check hooks
                          (273-292,298-308,573-577,592-593,605-611) left out: they only
                          assert/print and make keygen panic beyond 4-constraint toys
                          
  generator.rs:34-38      the fork's fixed toxic waste alpha=6 beta=24 gamma=6 delta=24 tau=2
  verifier.rs:10-62       prepare_verifying_key / verify_proof (DummyEngine: Gt = Fr, additive;
                          BLS12-381: the optimal ate pairing of oracle/pairing.py)

Engines: BLS12-381 (oracle.curves.G1/G2 over oracle.fields.Fr) and the reference's
DummyEngine (groth16/tests/dummy_engine.rs) so the golden vectors in
groth16/tests/mod.rs:334-435,574 pin this restatement.
"""
from __future__ import annotations

from . import curves, fields, pairing as _pairing
from .domain import EvaluationDomain
from .multiexp import (DensityTracker, FullDensity, SynthesisError, UnexpectedIdentity,
                       multiexp)


class Engine:
    """pairing::{Engine, MultiMillerLoop}: `pairing(p, q)` and `multi_miller_final(pairs)` =
    `multi_miller_loop(pairs).final_exponentiation()`; Gt values are compared with ==."""

    def __init__(self, name, Fr, G1, G2, pairing=None, multi_miller_final=None):
        self.name, self.Fr, self.G1, self.G2, self.pairing = name, Fr, G1, G2, pairing
        self.multi_miller_final = multi_miller_final


BLS12 = Engine("Bls12", fields.Fr, curves.G1, curves.G2, pairing=_pairing.pairing,
               multi_miller_final=lambda pairs: _pairing.final_exponentiation(_pairing.multi_miller_loop(pairs)))
# dummy_engine.rs:344-364: pairing(p, q) = p * q in Fr; Gt "multiplication" = addition, the Miller
# loop result is the sum of the products and the final exponentiation the identity map
DUMMY = Engine("DummyEngine", fields.DummyFr, curves.Dummy, curves.Dummy,
               pairing=lambda p, q: p * q % 64513,
               multi_miller_final=lambda pairs: sum(p * q for p, q in pairs) % 64513)

ONE = ("input", 0)   # ConstraintSystem::one()  (lib.rs)


class AssignmentMissing(SynthesisError):
    pass


class UnconstrainedVariable(SynthesisError):
    pass


# ----------------------------------------------------------------------- key generator
class KeypairAssembly:
    """generator.rs:44-156 -- column-major QAP: per variable a list of (coeff, constraint)."""

    def __init__(self):
        self.num_inputs = self.num_aux = self.num_constraints = 0
        self.at_inputs, self.bt_inputs, self.ct_inputs = [], [], []
        self.at_aux, self.bt_aux, self.ct_aux = [], [], []

    def alloc(self, f=None):
        self.num_aux += 1
        self.at_aux.append([]); self.bt_aux.append([]); self.ct_aux.append([])
        return ("aux", self.num_aux - 1)

    def alloc_input(self, f=None):
        self.num_inputs += 1
        self.at_inputs.append([]); self.bt_inputs.append([]); self.ct_inputs.append([])
        return ("input", self.num_inputs - 1)

    def enforce(self, a, b, c):
        def ev(lc, inputs, aux):
            for (kind, idx), coeff in lc:
                (inputs if kind == "input" else aux)[idx].append((coeff, self.num_constraints))
        ev(a, self.at_inputs, self.at_aux)
        ev(b, self.bt_inputs, self.bt_aux)
        ev(c, self.ct_inputs, self.ct_aux)
        self.num_constraints += 1


class VerifyingKey:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class Parameters:
    """mod.rs:224-247."""

    def __init__(self, vk, h, l, a, b_g1, b_g2):
        self.vk, self.h, self.l, self.a, self.b_g1, self.b_g2 = vk, h, l, a, b_g1, b_g2


def generate_parameters(engine, circuit, g1, g2, alpha, beta, gamma, delta, tau):
    """generator.rs:241-634, upstream semantics (see module docstring)."""
    F, G1, G2 = engine.Fr, engine.G1, engine.G2
    p = F.p
    asm = KeypairAssembly()
    asm.alloc_input()                                            # :263 the "one" input
    circuit(asm)                                                 # :266
    for i in range(asm.num_inputs):                              # :273-275 x * 0 = 0
        asm.enforce([(("input", i), 1)], [], [])

    dom = EvaluationDomain(F, [0] * asm.num_constraints)          # :294-297
    m = len(dom.coeffs)
    if gamma % p == 0 or delta % p == 0:                         # :330-345
        raise UnexpectedIdentity()
    gamma_inv, delta_inv = F.inv(gamma), F.inv(delta)

    powers = [pow(tau, i, p) for i in range(m)]                  # :351-366
    coeff = dom.z(tau) * delta_inv % p                           # :368-369
    h = [G1.mul(g1, powers[i] * coeff % p) for i in range(m - 1)]  # :372-397

    dom.coeffs = powers
    dom.ifft()                                                   # :401 Lagrange basis
    lag = dom.coeffs

    def eval_at_tau(poly):                                       # :471-484
        return sum(lag[idx] * c for c, idx in poly) % p

    def ev(at, bt, ct, inv):                                     # :419-536
        a_out, b1_out, b2_out, ext_out = [], [], [], []
        for ai, bi, ci in zip(at, bt, ct):
            atv, btv, ctv = eval_at_tau(ai), eval_at_tau(bi), eval_at_tau(ci)
            a_out.append(G1.mul(g1, atv) if atv else G1.identity())       # :492-494
            b1_out.append(G1.mul(g1, btv) if btv else G1.identity())      # :497-500
            b2_out.append(G2.mul(g2, btv) if btv else G2.identity())
            e = (atv * beta + btv * alpha + ctv) % p * inv % p              # :502-510
            ext_out.append(G1.mul(g1, e))
        return a_out, b1_out, b2_out, ext_out

    a_in, b1_in, b2_in, ic = ev(asm.at_inputs, asm.bt_inputs, asm.ct_inputs, gamma_inv)  # :539-554
    a_ax, b1_ax, b2_ax, l = ev(asm.at_aux, asm.bt_aux, asm.ct_aux, delta_inv)            # :557-572
    for e in l:                                                   # :584-590
        if G1.is_identity(e):
            raise UnconstrainedVariable()
    vk = VerifyingKey(alpha_g1=G1.mul(g1, alpha), beta_g1=G1.mul(g1, beta),
                      beta_g2=G2.mul(g2, beta), gamma_g2=G2.mul(g2, gamma),
                      delta_g1=G1.mul(g1, delta), delta_g2=G2.mul(g2, delta), ic=ic)
    keep1 = lambda v: [e for e in v if not G1.is_identity(e)]     # :618-632
    keep2 = lambda v: [e for e in v if not G2.is_identity(e)]
    params = Parameters(vk, h, l, keep1(a_in + a_ax), keep1(b1_in + b1_ax), keep2(b2_in + b2_ax))
    # the trapdoor-side scalars, kept for `expected_proof` (not part of the reference struct)
    params.qap = asm
    params.lagrange = lag
    params.trapdoor = dict(alpha=alpha, beta=beta, gamma=gamma, delta=delta, tau=tau, m=m)
    return params


def generate_random_parameters(engine, circuit):
    """generator.rs:21-40: RNG ignored, fixed toxic waste."""
    return generate_parameters(engine, circuit, engine.G1.gen, engine.G2.gen, 6, 24, 6, 24, 2)


# ----------------------------------------------------------------------------- prover
class ProvingAssignment:
    """prover.rs:55-156."""

    def __init__(self, F):
        self.F = F
        self.a_aux_density = DensityTracker()
        self.b_input_density = DensityTracker()
        self.b_aux_density = DensityTracker()
        self.a, self.b, self.c = [], [], []
        self.input_assignment, self.aux_assignment = [], []

    def alloc(self, f):
        v = f()
        if v is None:
            raise AssignmentMissing()
        self.aux_assignment.append(v % self.F.p)
        self.a_aux_density.add_element()
        self.b_aux_density.add_element()
        return ("aux", len(self.aux_assignment) - 1)

    def alloc_input(self, f):
        v = f()
        if v is None:
            raise AssignmentMissing()
        self.input_assignment.append(v % self.F.p)
        self.b_input_density.add_element()
        return ("input", len(self.input_assignment) - 1)

    def _eval(self, lc, input_density, aux_density):             # prover.rs:19-53
        acc = 0
        for (kind, idx), coeff in lc:
            if kind == "input":
                tmp = self.input_assignment[idx]
                if input_density is not None:
                    input_density.inc(idx)
            else:
                tmp = self.aux_assignment[idx]
                if aux_density is not None:
                    aux_density.inc(idx)
            acc = (acc + tmp * coeff) % self.F.p
        return acc

    def enforce(self, a, b, c):                                  # prover.rs:100-138
        self.a.append(self._eval(a, None, self.a_aux_density))
        self.b.append(self._eval(b, self.b_input_density, self.b_aux_density))
        self.c.append(self._eval(c, None, None))


class Proof:
    def __init__(self, a, b, c):
        self.a, self.b, self.c = a, b, c

    def to_bytes(self, engine):                                  # mod.rs:42-48 (192 B for BLS12)
        return (engine.G1.to_compressed(self.a) + engine.G2.to_compressed(self.b)
                + engine.G1.to_compressed(self.c))


    @staticmethod
    def read(engine, data):
        """mod.rs:50-103: three compressed points; "invalid G1/G2" for an undecodable one, "point at
        infinity" for the identity, UnexpectedEof for a short buffer (ValueError carries the message)."""
        n1, n2 = engine.G1.coord_bytes, engine.G2.coord_bytes
        if len(data) < 2 * n1 + n2:
            raise ValueError("UnexpectedEof")
        pts, off = [], 0
        for G, n, name in ((engine.G1, n1, "G1"), (engine.G2, n2, "G2"), (engine.G1, n1, "G1")):
            pt = G.from_compressed(bytes(data[off:off + n]))
            off += n
            if pt is None:
                raise ValueError("invalid " + name)
            if pt == "identity":
                raise ValueError("point at infinity")
            pts.append(pt)
        return Proof(*pts)


def synthesize_for_proving(engine, circuit):
    """prover.rs:187-204: alloc ONE, synthesize, one `x * 0 = 0` per input."""
    prover = ProvingAssignment(engine.Fr)
    prover.alloc_input(lambda: 1)
    circuit(prover)
    for i in range(len(prover.input_assignment)):
        prover.enforce([(("input", i), 1)], [], [])
    return prover


def h_coefficients(F, a, b, c):
    """prover.rs:210-229: the quotient coefficients fed to the H multiexp (m-1 of them)."""
    A, B, C = EvaluationDomain(F, a), EvaluationDomain(F, b), EvaluationDomain(F, c)
    for d in (A, B, C):
        d.ifft()
        d.coset_fft()
    A.mul_assign(B)
    A.sub_assign(C)
    A.divide_by_z_on_coset()
    A.icoset_fft()
    out = A.into_coeffs()
    return out[:-1]


def create_proof_from_assignment(engine, prover, params, r, s):
    """prover.rs:206-350 -- everything after synthesis."""
    F, G1, G2 = engine.Fr, engine.G1, engine.G2
    vk = params.vk
    h_exps = h_coefficients(F, prover.a, prover.b, prover.c)
    h = multiexp(G1, params.h, 0, FullDensity(), h_exps)                          # :233
    inputs, aux = prover.input_assignment, prover.aux_assignment
    l = multiexp(G1, params.l, 0, FullDensity(), aux)                              # :252-257
    a_inputs = multiexp(G1, params.a, 0, FullDensity(), inputs)                    # :264-269
    a_aux = multiexp(G1, params.a, len(inputs), prover.a_aux_density, aux)         # :270-275
    b_in_tot = prover.b_input_density.get_total_density()
    b_g1_inputs = multiexp(G1, params.b_g1, 0, prover.b_input_density, inputs)     # :285-296
    b_g1_aux = multiexp(G1, params.b_g1, b_in_tot, prover.b_aux_density, aux)
    b_g2_inputs = multiexp(G2, params.b_g2, 0, prover.b_input_density, inputs)     # :301-307
    b_g2_aux = multiexp(G2, params.b_g2, b_in_tot, prover.b_aux_density, aux)
    if G1.is_identity(vk.delta_g1) or G2.is_identity(vk.delta_g2):                 # :309-313
        raise UnexpectedIdentity()
    p = F.p
    g_a = G1.add(G1.mul(vk.delta_g1, r), vk.alpha_g1)                              # :315-316
    g_b = G2.add(G2.mul(vk.delta_g2, s), vk.beta_g2)                               # :317-318
    g_c = G1.mul(vk.delta_g1, r * s % p)                                           # :319-327
    g_c = G1.add(g_c, G1.mul(vk.alpha_g1, s))
    g_c = G1.add(g_c, G1.mul(vk.beta_g1, r))
    a_answer = G1.add(a_inputs, a_aux)                                             # :328-332
    g_a = G1.add(g_a, a_answer)
    g_c = G1.add(g_c, G1.mul(a_answer, s))
    b1_answer = G1.add(b_g1_inputs, b_g1_aux)                                      # :334-343
    b2_answer = G2.add(b_g2_inputs, b_g2_aux)
    g_b = G2.add(g_b, b2_answer)
    g_c = G1.add(g_c, G1.mul(b1_answer, r))
    g_c = G1.add(g_c, h)
    g_c = G1.add(g_c, l)
    return Proof(g_a, g_b, g_c)


def create_proof(engine, circuit, params, r, s):
    return create_proof_from_assignment(engine, synthesize_for_proving(engine, circuit), params, r, s)


def create_random_proof(engine, circuit, params):
    """prover.rs:158-173."""
    return create_proof(engine, circuit, params, 27134, 17146)


# -------------------------------------------------------------------------- verifier
class PreparedVerifyingKey:
    """groth16/mod.rs:403-414"""

    def __init__(self, alpha_g1_beta_g2, neg_gamma_g2, neg_delta_g2, ic):
        self.alpha_g1_beta_g2, self.neg_gamma_g2, self.neg_delta_g2, self.ic = (
            alpha_g1_beta_g2, neg_gamma_g2, neg_delta_g2, ic)


def prepare_verifying_key(engine, vk):
    """verifier.rs:10-21"""
    G2 = engine.G2
    return PreparedVerifyingKey(engine.pairing(vk.alpha_g1, vk.beta_g2), G2.neg(vk.gamma_g2),
                                G2.neg(vk.delta_g2), list(vk.ic))


class InvalidVerifyingKey(Exception):
    """VerificationError::InvalidVerifyingKey (lib.rs)"""


def verify_proof_prepared(engine, pvk, proof, public_inputs):
    """verifier.rs:23-62: Err(InvalidVerifyingKey) raises, Err(InvalidProof) -> False, Ok -> True."""
    G1 = engine.G1
    if len(public_inputs) + 1 != len(pvk.ic):
        raise InvalidVerifyingKey()
    acc = pvk.ic[0]
    for i, b in zip(public_inputs, pvk.ic[1:]):
        acc = G1.add(acc, G1.mul(b, i))
    # A * B + inputs * (-gamma) + C * (-delta) = alpha * beta with a single final exponentiation
    return pvk.alpha_g1_beta_g2 == engine.multi_miller_final(
        [(proof.a, proof.b), (acc, pvk.neg_gamma_g2), (proof.c, pvk.neg_delta_g2)])


def verify_proof(engine, vk, proof, public_inputs):
    """prepare_verifying_key + verify_proof as the reference's tests chain them
    (groth16/tests/mod.rs:574-585, tests/mimc.rs); a wrong input count returns False."""
    try:
        return verify_proof_prepared(engine, prepare_verifying_key(engine, vk), proof, public_inputs)
    except InvalidVerifyingKey:
        return False


# ------------------------------------------------------------- known-trapdoor expectation
def expected_proof(engine, params, prover, r, s):
    """with the trapdoor known, the proof is computable in the scalar
    field -- A = g1*(alpha + sum w_i u_i(tau) + r delta) etc. -- which pins the whole
    create_proof pipeline (NTTs, densities, offsets, MSMs, tail) without a CPU MSM."""
    F, G1, G2 = engine.Fr, engine.G1, engine.G2
    p = F.p
    t = params.trapdoor
    alpha, beta, delta, tau, m = t["alpha"], t["beta"], t["delta"], t["tau"], t["m"]
    lag, asm = params.lagrange, params.qap
    w = prover.input_assignment + prover.aux_assignment

    def at_tau(cols_in, cols_aux):
        return [sum(lag[idx] * c for c, idx in col) % p for col in cols_in + cols_aux]

    u, v, wq = (at_tau(asm.at_inputs, asm.at_aux), at_tau(asm.bt_inputs, asm.bt_aux),
                at_tau(asm.ct_inputs, asm.ct_aux))
    sa = sum(wi * ui for wi, ui in zip(w, u)) % p
    sb = sum(wi * vi for wi, vi in zip(w, v)) % p
    h = h_coefficients(F, prover.a, prover.b, prover.c)
    ttau = (pow(tau, m, p) - 1) % p
    dinv = F.inv(delta)
    hs = sum(hi * pow(tau, i, p) for i, hi in enumerate(h)) % p * ttau % p * dinv % p
    ni = len(prover.input_assignment)
    ls = sum(w[i] * (beta * u[i] + alpha * v[i] + wq[i]) for i in range(ni, len(w))) % p * dinv % p
    a_s = (alpha + sa + r * delta) % p
    b_s = (beta + sb + s * delta) % p
    c_s = (r * s % p * delta + s * alpha + r * beta + s * sa + r * sb + hs + ls) % p
    return Proof(G1.mul(G1.gen, a_s), G2.mul(G2.gen, b_s), G1.mul(G1.gen, c_s))


# ---------------------------------------------------------------------- demo circuits
def xor_demo(a, b):
    """groth16/tests/mod.rs:86-161 XorDemo."""
    def synth(cs):
        a_var = cs.alloc(lambda: None if a is None else int(a))
        cs.enforce([(ONE, 1), (a_var, -1)], [(a_var, 1)], [])
        b_var = cs.alloc(lambda: None if b is None else int(b))
        cs.enforce([(ONE, 1), (b_var, -1)], [(b_var, 1)], [])
        c_var = cs.alloc_input(lambda: None if a is None or b is None else int(a ^ b))
        cs.enforce([(a_var, 1), (a_var, 1)], [(b_var, 1)], [(a_var, 1), (b_var, 1), (c_var, -1)])
    return synth


MIMC_ROUNDS = 322


def mimc(F, xl, xr, constants):
    """mimc_mod.rs:22-37."""
    p = F.p
    for c in constants:
        t = (xl + c) % p
        xl, xr = (t * t % p * t + xr) % p, xl
    return xl


def mimc_demo(F, xl, xr, constants):
    """mimc_mod.rs:49-129 MiMCDemo::synthesize."""
    p = F.p

    def synth(cs):
        xl_v, xr_v = xl, xr
        xl_var = cs.alloc(lambda: xl_v)
        xr_var = cs.alloc(lambda: xr_v)
        for i in range(MIMC_ROUNDS):
            ci = constants[i]
            tmp_v = None if xl_v is None else (xl_v + ci) % p * ((xl_v + ci) % p) % p
            tmp = cs.alloc(lambda: tmp_v)
            cs.enforce([(xl_var, 1), (ONE, ci)], [(xl_var, 1), (ONE, ci)], [(tmp, 1)])
            new_v = None if xl_v is None else ((xl_v + ci) % p * tmp_v + xr_v) % p
            if i == MIMC_ROUNDS - 1:
                new_xl = cs.alloc_input(lambda: new_v)
            else:
                new_xl = cs.alloc(lambda: new_v)
            cs.enforce([(tmp, 1)], [(xl_var, 1), (ONE, ci)], [(new_xl, 1), (xr_var, -1)])
            xr_var, xr_v = xl_var, xl_v
            xl_var, xl_v = new_xl, new_v
    return synth
