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
