"""TEST INFRASTRUCTURE ONLY -- restatement of the ceremony's point-vector operations.

Follows /root/reference/bellman/src/groth16/mpc.rs:
  list_mul_matrix ................................. mpc.rs:416-457
  make_new_paramter / make_new_tau_paramter ....... mpc.rs:647-706 (element * scalar)

Generic over any group object from oracle.curves (G1, G2, Dummy), as the reference is generic
over `Engine`.  Parity pinning: the reference holds no known-answer vector for these functions
(mpc.rs has no tests); results are canonical group elements, and the tests pin this restatement
through bases with known discrete logarithms: result[i] == (sum_j m_ij * k_idx_j) * G.
"""
from __future__ import annotations


def list_mul_matrix(group, lst, matrix):
    """mpc.rs:416-457 for one of the two lists (the reference runs the identical loop on a G1
    and a G2 list).  `matrix[i]` = [(coefficient, index), ...]."""
    n = len(lst)
    result = [group.identity() for _ in range(n)]            # :428-429
    for i in range(len(matrix)):                             # :430
        if len(matrix[i]) == 0:                              # :432-434  `break`, not `continue`
            break
        for (coeff, idx) in matrix[i]:                       # :444-447
            if i >= n or idx >= n:
                raise IndexError("index out of bounds")      # Rust slice indexing panics
            result[i] = group.add(result[i], group.mul(lst[idx], coeff))
    return result


def make_new_parameter(group, lst, scalars):
    """mpc.rs:647-706: every element times its own scalar."""
    return [group.mul(p, k) for p, k in zip(lst, scalars)]


# ------------------------------------------------------------------ contribution checks (pairings)
class ParameterPair:
    """mpc.rs:19-29: a participant's update of one CRS element: the new element in both groups and the
    participant's own factor in both groups."""

    def __init__(self, g1_result, g2_result, g1_mine, g2_mine):
        self.g1_result, self.g2_result, self.g1_mine, self.g2_mine = g1_result, g2_result, g1_mine, g2_mine


def verify_new_parameter(engine, pair, base_g1, base_g2):
    """mpc.rs:787-804 `verify_new_paramter`: the new element is the stored one times the
    participant's factor, and its G1 and G2 forms agree.  (`base_g2` is unused by the reference too.)"""
    e, G1, G2 = engine.pairing, engine.G1, engine.G2
    lhs = e(pair.g1_result, G2.gen)
    return lhs == e(base_g1, pair.g2_mine) and lhs == e(G1.gen, pair.g2_result)


def verify_vector(engine, new_list_g1, g2_point, matrixed_g1):
    """The three per-element loops of `verify_uncommon_paramter` (mpc.rs:1091-1124: ic against gamma,
    l and h against delta): for every i, e(new[i], g2_point) == e(matrixed[i], G2::generator())."""
    e, G2 = engine.pairing, engine.G2
    ok = True
    for new, old in zip(new_list_g1, matrixed_g1):
        ok = ok and e(new, g2_point) == e(old, G2.gen)
    return ok


def verify_vector_folded(engine, folded_new, g2_point, folded_matrixed):
    """The same n equations checked at once on a random linear combination: with
    N = sum rho_i new[i] and M = sum rho_i matrixed[i] (the two multiexps the product computes on the
    GPU), e(N, g2_point) == e(M, G2::generator()).  Equivalent to `verify_vector` except with
    probability 2^-128 over 128-bit rho (a cheating element survives only if it cancels in the
    combination).  This is a semantic change relative to the reference's per-element loop,
    which is why it is offered beside, not instead of, the per-element check."""
    e, G2 = engine.pairing, engine.G2
    return e(folded_new, g2_point) == e(folded_matrixed, G2.gen)
