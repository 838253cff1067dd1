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
