"""The device arithmetic headers (field.cuh / curve.cuh) compiled for the host with the PTX
carry flag emulated, checked bit-for-bit against the big-integer oracle: Montgomery products,
XYZZ group law incl. the exceptional cases."""
import ctypes
import os
import random
import subprocess

import pytest

from oracle import curves, fields

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "host_check", "libhost_check.so")


@pytest.fixture(scope="module")
def lib():
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", SO,
                    os.path.join(HERE, "host_check", "host_check.cpp")], check=True)
    return ctypes.CDLL(SO)


def _l(x, n):
    return (ctypes.c_uint32 * n)(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)])


def _v(arr):
    return sum(int(v) << (32 * i) for i, v in enumerate(arr))


@pytest.mark.parametrize("which", ["fr", "fp"])
def test_field_ops(lib, which):
    F, fn, n = (fields.Fr, lib.hc_fr_op, 8) if which == "fr" else (fields.Fp, lib.hc_fp_op, 12)
    p, R = F.p, F.R
    Ri = pow(R, -1, p)
    rng = random.Random(1)
    cases = [(0, 0), (1, 1), (p - 1, p - 1), (p - 1, 1), (R, R)] + [(rng.randrange(p), rng.randrange(p)) for _ in range(500)]
    for a, b in cases:
        out = (ctypes.c_uint32 * n)()
        for op, e in {0: a * b * Ri % p, 1: (a + b) % p, 2: (a - b) % p, 3: (-a) % p, 5: a * R % p, 6: a * Ri % p, 7: a * a * Ri % p}.items():
            fn(op, _l(a, n), _l(b, n), out)
            assert _v(out) == e, (which, op)
    for a, _ in cases[1:20]:
        out = (ctypes.c_uint32 * n)()
        fn(4, _l(a, n), _l(0, n), out)
        assert _v(out) == R * R * pow(a, -1, p) % p


def test_group_law(lib):
    P, R = fields.Fp.p, fields.Fp.R
    Ri = pow(R, -1, P)
    fl = lambda x: [((x * R % P) >> (32 * i)) & 0xFFFFFFFF for i in range(12)]
    fv = lambda l: _v(l) * Ri % P

    def enc1(p):
        return (ctypes.c_uint32 * 24)() if p is None else (ctypes.c_uint32 * 24)(*(fl(p[0]) + fl(p[1])))

    def dec1(a):
        a = list(a)
        return None if not any(a) else (fv(a[:12]), fv(a[12:]))

    def enc2(p):
        if p is None:
            return (ctypes.c_uint32 * 48)()
        return (ctypes.c_uint32 * 48)(*(fl(p[0][0]) + fl(p[0][1]) + fl(p[1][0]) + fl(p[1][1])))

    def dec2(a):
        a = list(a)
        return None if not any(a) else ((fv(a[:12]), fv(a[12:24])), (fv(a[24:36]), fv(a[36:])))

    kl = lambda k: (ctypes.c_uint32 * 8)(*[(k >> (32 * i)) & 0xFFFFFFFF for i in range(8)])
    rng = random.Random(2)
    for G, fn, enc, dec, n in ((curves.G1, lib.hc_g1_op, enc1, dec1, 24), (curves.G2, lib.hc_g2_op, enc2, dec2, 48)):
        pts = [G.mul(G.gen, rng.randrange(1, fields.Fr.p)) for _ in range(5)]
        cases = [(pts[0], pts[1]), (pts[0], pts[0]), (pts[0], G.neg(pts[0])), (None, pts[2]), (pts[2], None), (None, None)]
        for p, q in cases:
            for op in (0, 1, 2):
                out = (ctypes.c_uint32 * n)()
                fn(op, enc(p), enc(q), kl(0), out)
                assert dec(out) == (G.add(p, q) if op < 2 else G.double(p)), (G.name, op)
        for k in (0, 1, 2, fields.Fr.p - 1, rng.randrange(fields.Fr.p)):
            out = (ctypes.c_uint32 * n)()
            fn(3, enc(pts[4]), enc(None), kl(k), out)
            assert dec(out) == G.mul(pts[4], k)
