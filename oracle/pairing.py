"""TEST INFRASTRUCTURE ONLY -- optimal ate pairing on BLS12-381 with plain Python integers.

The reference verifies a proof with `E::multi_miller_loop(&[(A, B), (acc, -gamma), (C, -delta)])
.final_exponentiation() == E::pairing(alpha, beta)` (src/groth16/verifier.rs:10-21,23-62) and checks
ceremony contributions with `E::pairing(..) == E::pairing(..)` (src/groth16/mpc.rs:164-177,793-836,
1091-1119).  The pairing itself lives in the un-vendored crate bls12_381 0.6.0 (Cargo.lock:96-99;
`pairing 0.21.0` supplies the traits), so this file restates the *published* construction:

  Fp2 = Fp[u]/(u^2 + 1),  Fp12 = Fp2[w]/(w^6 - xi),  xi = 1 + u          (w^2 = v, v^3 = xi: the
                                                                          crate's Fp6/Fp12 tower)
  E : y^2 = x^3 + 4 over Fp,  E' : y^2 = x^3 + 4 xi over Fp2 (M-type twist),
  untwist psi(x', y') = (x' / w^2, y' / w^3)
  e(P, Q) = f_{|x|, Q}(P)^((p^12 - 1) / r), conjugated because the BLS parameter x = -0xd201000000010000
  is negative.

Every use the reference makes of a pairing value is an equality test between products of pairings,
which holds or fails identically for any non-degenerate bilinear map on (G1, G2) -- in particular for
any fixed power of the pairing coprime to r (libraries differ by such powers in the hard part of the
final exponentiation).  `self_check()` pins exactly the properties those tests rely on: values have
order r, are not 1, and e(aP, bQ) = e(P, Q)^(ab).

Elements of Fp12 are 6-tuples of Fp2 pairs: a = sum a[i] w^i.  Slow (tenths of a second per pairing)
and only meant for the handful of proofs the tests verify.
"""
from __future__ import annotations

from .fields import FP_MODULUS, FR_MODULUS
from . import curves
from .curves import fp2_add, fp2_sub, fp2_mul, fp2_neg, fp2_inv, fp2_sqr

P = FP_MODULUS
R = FR_MODULUS
BLS_X = 0xD201000000010000          # |x|; the curve parameter is -|x|
XI = (1, 1)
F2_ZERO, F2_ONE = (0, 0), (1, 0)
F12_ONE = (F2_ONE,) + (F2_ZERO,) * 5


def fp2_mul_xi(a):
    """a (1 + u)"""
    return ((a[0] - a[1]) % P, (a[0] + a[1]) % P)


def fp2_conj(a):
    return (a[0], (-a[1]) % P)


def fp2_pow(a, e):
    r = F2_ONE
    while e:
        if e & 1:
            r = fp2_mul(r, a)
        a = fp2_sqr(a)
        e >>= 1
    return r


# ----------------------------------------------------------------------------- Fp12 = Fp2[w]/(w^6 - xi)
def f12_mul(a, b):
    t = [F2_ZERO] * 11
    for i in range(6):
        ai = a[i]
        if ai == F2_ZERO:
            continue
        for j in range(6):
            bj = b[j]
            if bj == F2_ZERO:
                continue
            t[i + j] = fp2_add(t[i + j], fp2_mul(ai, bj))
    return tuple(fp2_add(t[i], fp2_mul_xi(t[i + 6])) if i < 5 else t[5] for i in range(6))


def f12_sqr(a):
    return f12_mul(a, a)


def f12_conj(a):
    """a^(p^6): w -> -w"""
    return tuple(a[i] if i % 2 == 0 else fp2_neg(a[i]) for i in range(6))


# Frobenius: (sum a_i w^i)^p = sum conj(a_i) gamma^i w^i with gamma = w^(p-1) = xi^((p-1)/6)
_GAMMA1 = fp2_pow(XI, (P - 1) // 6)
_GAMMA = [fp2_pow(_GAMMA1, i) for i in range(6)]


def f12_frob(a):
    return tuple(fp2_mul(fp2_conj(a[i]), _GAMMA[i]) for i in range(6))


# inversion through the tower Fp12 = Fp6[w]/(w^2 - v), Fp6 = Fp2[v]/(v^3 - xi):
# a = (a0 + a2 v + a4 v^2) + (a1 + a3 v + a5 v^2) w
def _f6_mul(a, b):
    a0, a1, a2 = a
    b0, b1, b2 = b
    t0, t1, t2 = fp2_mul(a0, b0), fp2_mul(a1, b1), fp2_mul(a2, b2)
    c0 = fp2_add(t0, fp2_mul_xi(fp2_add(fp2_mul(a1, b2), fp2_mul(a2, b1))))
    c1 = fp2_add(fp2_add(fp2_mul(a0, b1), fp2_mul(a1, b0)), fp2_mul_xi(t2))
    c2 = fp2_add(fp2_add(fp2_mul(a0, b2), fp2_mul(a2, b0)), t1)
    return (c0, c1, c2)


def _f6_mul_v(a):
    return (fp2_mul_xi(a[2]), a[0], a[1])


def _f6_sub(a, b):
    return tuple(fp2_sub(x, y) for x, y in zip(a, b))


def _f6_inv(a):
    a0, a1, a2 = a
    c0 = fp2_sub(fp2_sqr(a0), fp2_mul_xi(fp2_mul(a1, a2)))
    c1 = fp2_sub(fp2_mul_xi(fp2_sqr(a2)), fp2_mul(a0, a1))
    c2 = fp2_sub(fp2_sqr(a1), fp2_mul(a0, a2))
    n = fp2_add(fp2_mul(a0, c0), fp2_mul_xi(fp2_add(fp2_mul(a2, c1), fp2_mul(a1, c2))))
    ni = fp2_inv(n)
    return (fp2_mul(c0, ni), fp2_mul(c1, ni), fp2_mul(c2, ni))


def f12_inv(a):
    c0, c1 = (a[0], a[2], a[4]), (a[1], a[3], a[5])
    n = _f6_sub(_f6_mul(c0, c0), _f6_mul_v(_f6_mul(c1, c1)))
    ni = _f6_inv(n)
    r0 = _f6_mul(c0, ni)
    r1 = _f6_mul(c1, ni)
    return (r0[0], fp2_neg(r1[0]), r0[1], fp2_neg(r1[1]), r0[2], fp2_neg(r1[2]))


def f12_pow(a, e):
    r = F12_ONE
    for bit in bin(e)[2:]:
        r = f12_sqr(r)
        if bit == "1":
            r = f12_mul(r, a)
    return r


# ----------------------------------------------------------------------------- Miller loop
def _line(T, lam, Pt):
    """The line through psi(T) with twist-slope lam, evaluated at P in E(Fp) and scaled by w^3 (an
    element of a proper subfield, killed by the final exponentiation):
    yP w^3 - lam xP w^2 + (lam xT - yT)."""
    xP, yP = Pt
    xT, yT = T
    c0 = fp2_sub(fp2_mul(lam, xT), yT)
    c2 = fp2_neg((lam[0] * xP % P, lam[1] * xP % P))
    c3 = (yP % P, 0)
    return (c0, F2_ZERO, c2, c3, F2_ZERO, F2_ZERO)


def miller_loop(Pt, Q):
    """f_{|x|, Q}(P) for P in G1 (affine over Fp), Q in G2 (affine on the twist over Fp2); identity
    operands give 1, as in bls12_381's multi_miller_loop."""
    if Pt is None or Q is None:
        return F12_ONE
    f = F12_ONE
    T = Q
    three = 3
    for bit in bin(BLS_X)[3:]:
        # doubling step
        xT, yT = T
        xx = fp2_sqr(xT)
        lam = fp2_mul((xx[0] * three % P, xx[1] * three % P), fp2_inv(fp2_add(yT, yT)))
        f = f12_mul(f12_sqr(f), _line(T, lam, Pt))
        x3 = fp2_sub(fp2_sub(fp2_sqr(lam), xT), xT)
        T = (x3, fp2_sub(fp2_mul(lam, fp2_sub(xT, x3)), yT))
        if bit == "1":
            xT, yT = T
            lam = fp2_mul(fp2_sub(Q[1], yT), fp2_inv(fp2_sub(Q[0], xT)))
            f = f12_mul(f, _line(T, lam, Pt))
            x3 = fp2_sub(fp2_sub(fp2_sqr(lam), xT), Q[0])
            T = (x3, fp2_sub(fp2_mul(lam, fp2_sub(xT, x3)), yT))
    return f12_conj(f)                  # negative x


def multi_miller_loop(pairs):
    """pairing::MultiMillerLoop::multi_miller_loop: the product of the Miller functions"""
    f = F12_ONE
    for Pt, Q in pairs:
        f = f12_mul(f, miller_loop(Pt, Q))
    return f


_HARD = (P ** 4 - P ** 2 + 1) // R
assert (P ** 4 - P ** 2 + 1) % R == 0


def final_exponentiation(f):
    """f^((p^12 - 1) / r) = ((f^(p^6 - 1))^(p^2 + 1))^((p^4 - p^2 + 1) / r)"""
    f = f12_mul(f12_conj(f), f12_inv(f))
    f = f12_mul(f12_frob(f12_frob(f)), f)
    return f12_pow(f, _HARD)


def pairing(Pt, Q):
    return final_exponentiation(miller_loop(Pt, Q))


def self_check():
    G1, G2 = curves.G1, curves.G2
    # tower constants
    assert fp2_pow(_GAMMA1, 6) == fp2_pow(XI, P - 1)
    a = tuple(((i * 7 + 3) % P, (i * 11 + 5) % P) for i in range(6))
    assert f12_mul(a, f12_inv(a)) == F12_ONE
    fa, ap = a, f12_pow(a, P)
    assert f12_frob(fa) == ap
    e = pairing(G1.gen, G2.gen)
    assert e != F12_ONE
    assert f12_pow(e, R) == F12_ONE
    # bilinearity in both arguments
    assert pairing(G1.mul(G1.gen, 5), G2.mul(G2.gen, 7)) == f12_pow(e, 35)
    assert pairing(G1.neg(G1.gen), G2.gen) == f12_inv(e) == pairing(G1.gen, G2.neg(G2.gen))
    assert pairing(None, G2.gen) == F12_ONE and pairing(G1.gen, None) == F12_ONE
    return True
