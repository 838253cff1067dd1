"""CPU: the oracle restatement of mpc.rs:416-457 (list_mul_matrix) pinned through known discrete
logarithms on the DummyEngine group (F_64513, exact integers -- dummy_engine.rs:15) and BLS12-381."""
import random

import pytest

from oracle import curves, fields, mpc


def _matrix(rng, rows, n, q, empty_at=None):
    m = []
    for i in range(rows):
        k = 0 if i == empty_at else rng.randrange(1, 5)
        m.append([(rng.choice([0, 1, q - 1, rng.randrange(q)]), rng.randrange(n)) for _ in range(k)])
    return m


def test_list_mul_matrix_dummy_group():
    G, q = curves.Dummy, fields.DummyFr.p
    rng = random.Random(5)
    n = 40
    lst = [rng.randrange(q) for _ in range(n)]            # dummy group: element == its own dlog
    m = _matrix(rng, 33, n, q, empty_at=20)
    got = mpc.list_mul_matrix(G, lst, m)
    assert len(got) == n
    for i in range(n):
        want = sum(cf * lst[idx] for cf, idx in m[i]) % q if i < 20 else 0
        assert got[i] == want
    assert any(m[i] for i in range(21, 33))               # rows after the break exist and are ignored


def test_list_mul_matrix_bls_known_dlog():
    q = fields.Fr.p
    rng = random.Random(6)
    for G in (curves.G1, curves.G2):
        ks = [rng.randrange(1, q) for _ in range(6)]
        lst = [G.mul(G.gen, k) for k in ks]
        lst[2] = G.identity()
        ks[2] = 0
        m = _matrix(rng, 5, 6, q)
        got = mpc.list_mul_matrix(G, lst, m)
        for i in range(6):
            dot = sum(cf * ks[idx] for cf, idx in m[i]) % q if i < 5 else 0
            assert G.to_uncompressed(got[i]) == G.to_uncompressed(G.mul(G.gen, dot))


def test_list_mul_matrix_index_panics():
    G, q = curves.Dummy, fields.DummyFr.p
    with pytest.raises(IndexError):
        mpc.list_mul_matrix(G, [1, 2], [[(1, 0)], [(1, 1)], [(1, 0)]])      # taller than the list
    with pytest.raises(IndexError):
        mpc.list_mul_matrix(G, [1, 2], [[(1, 2)]])                           # column out of range


def test_contribution_checks_dummy_and_bls():
    """verify_new_paramter (mpc.rs:787-804) and the per-element loops of verify_uncommon_paramter
    (:1091-1124) with the oracle pairings; the folded form accepts and rejects the same cases."""
    from oracle import groth16 as og
    rng = random.Random(12)
    for E in (og.DUMMY, og.BLS12):
        q = E.Fr.p
        G1, G2 = E.G1, E.G2
        base, mine = rng.randrange(1, q), rng.randrange(1, q)
        pair = mpc.ParameterPair(G1.mul(G1.gen, base * mine % q), G2.mul(G2.gen, base * mine % q),
                                 G1.mul(G1.gen, mine), G2.mul(G2.gen, mine))
        assert mpc.verify_new_parameter(E, pair, G1.mul(G1.gen, base), G2.mul(G2.gen, base))
        bad = mpc.ParameterPair(pair.g1_result, G2.mul(G2.gen, (base * mine + 1) % q), pair.g1_mine, pair.g2_mine)
        assert not mpc.verify_new_parameter(E, bad, G1.mul(G1.gen, base), G2.mul(G2.gen, base))
        # l against delta: new[i] = matrixed[i] / delta
        n = 3
        delta = rng.randrange(1, q)
        ms = [rng.randrange(1, q) for _ in range(n)]
        matrixed = [G1.mul(G1.gen, m) for m in ms]
        new = [G1.mul(G1.gen, m * pow(delta, -1, q) % q) for m in ms]
        d2 = G2.mul(G2.gen, delta)
        assert mpc.verify_vector(E, new, d2, matrixed)
        rho = [rng.randrange(1 << 128) % q for _ in range(n)]
        fold = lambda pts: G1.sum([G1.mul(p, r) for p, r in zip(pts, rho)]) if E is og.BLS12 else \
            sum(p * r for p, r in zip(pts, rho)) % q
        assert mpc.verify_vector_folded(E, fold(new), d2, fold(matrixed))
        new[1] = G1.add(new[1], G1.gen)
        assert not mpc.verify_vector(E, new, d2, matrixed)
        assert not mpc.verify_vector_folded(E, fold(new), d2, fold(matrixed))
