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


@pytest.mark.parametrize("which", ["fr", "fp"])
def test_inversion(lib, which):
    """Field::inv (Kaliski almost-inverse + two Montgomery products, field.cuh) against pow(a, -1, p):
    small values (k at its minimum: the doubling fix-up), powers of two and values with long runs of zero
    bits (multi-bit shifts, a zero low limb), p - 1, R, R^-1, and random elements; inv(0) = 0."""
    F, fn, n = (fields.Fr, lib.hc_fr_op, 8) if which == "fr" else (fields.Fp, lib.hc_fp_op, 12)
    p, R = F.p, F.R
    rng = random.Random(5)
    cases = [1, 2, 3, p - 1, p - 2, R % p, R * R % p, pow(R, -1, p), 1 << 31, 1 << 32, 1 << 64, 1 << 96,
             1 << (p.bit_length() - 1), p >> 1, (p >> 1) + 1]
    cases += [rng.randrange(1, p) for _ in range(1500)]
    cases += [rng.randrange(1, 1 << 40) for _ in range(100)]
    cases += [(rng.randrange(1, p) >> rng.randrange(1, 200)) << rng.randrange(0, 100) for _ in range(200)]
    for a in cases:
        a %= p
        if not a:
            continue
        out = (ctypes.c_uint32 * n)()
        fn(4, _l(a, n), _l(0, n), out)
        assert _v(out) == R * R * pow(a, -1, p) % p, (which, hex(a))
    out = (ctypes.c_uint32 * n)()
    fn(4, _l(0, n), _l(0, n), out)
    assert _v(out) == 0


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


@pytest.mark.parametrize("which", ["g1", "g2"])
def test_batched_affine_job(lib, which):
    """msm_affine.cuh's per-thread job (pairwise tree, Montgomery's trick over the chord/tangent
    denominators) == the XYZZ chain == the oracle's sum, incl. duplicates, opposite points,
    identity bases, slices of length 0/1/2/3/odd/long (several inversion chunks)."""
    P, R = fields.Fp.p, fields.Fp.R
    Ri = pow(R, -1, P)
    fl = lambda x: [((x * R % P) >> (32 * i)) & 0xFFFFFFFF for i in range(12)]
    fv = lambda l: _v(l) * Ri % P
    rng = random.Random(5)
    if which == "g1":
        G, fn, words = curves.G1, lib.hc_affine_job_g1, 24
        enc = lambda p: [0] * 24 if p is None else fl(p[0]) + fl(p[1])
        dec = lambda a: None if not any(a) else (fv(a[:12]), fv(a[12:]))
    else:
        G, fn, words = curves.G2, lib.hc_affine_job_g2, 48
        enc = lambda p: [0] * 48 if p is None else fl(p[0][0]) + fl(p[0][1]) + fl(p[1][0]) + fl(p[1][1])
        dec = lambda a: None if not any(a) else ((fv(a[:12]), fv(a[12:24])), (fv(a[24:36]), fv(a[36:])))
    ng = lib.hc_affine_g()
    nb = 24
    pts = [G.mul(G.gen, rng.randrange(1, fields.Fr.p)) for _ in range(nb - 1)] + [None]
    pts[5] = pts[4]                       # duplicate base
    bases = (ctypes.c_uint32 * (words * nb))(*sum((enc(p) for p in pts), []))
    L = 300
    lens = ([0, 1, 2, 3, 5, 8, 13, 300] + [rng.randrange(0, 40) for _ in range(ng)])[:ng]
    if ng < 8:
        lens[-1] = 300
    sorted_entries, tasks = [], []
    for ln in lens:
        tasks += [len(sorted_entries), ln]
        for j in range(ln):
            r = rng.random()
            if r < 0.15 and j > 0:
                e = sorted_entries[-1] ^ 0x80000000          # the negative of the previous point
            elif r < 0.3 and j > 0:
                e = sorted_entries[-1]                       # the same point again (tangent case)
            else:
                e = rng.randrange(nb) | (rng.randrange(2) << 31)
            sorted_entries.append(e)
    # a slice made only of cancelling pairs and identities
    srt = (ctypes.c_uint32 * max(1, len(sorted_entries)))(*sorted_entries)
    tk = (ctypes.c_uint32 * (2 * ng))(*tasks)
    out_tree = (ctypes.c_uint32 * (words * ng))()
    out_chain = (ctypes.c_uint32 * (words * ng))()
    fn(bases, srt, tk, L, out_tree, out_chain)
    tree, chain = list(out_tree), list(out_chain)
    assert tree == chain
    for g in range(ng):
        expect = None
        for e in sorted_entries[tasks[2 * g]:tasks[2 * g] + tasks[2 * g + 1]]:
            p = pts[e & 0x7FFFFFFF]
            expect = G.add(expect, G.neg(p) if e >> 31 else p)
        assert dec(tree[g * words:(g + 1) * words]) == expect, (which, g)


@pytest.mark.parametrize("which", ["g1", "g2"])
def test_pair_rounds(lib, which):
    """msm_pairs.cuh on the host: the per-slice list rule (pair_build_task: round-0 pairs straight from
    the even-padded entry array, explicit lists with a CARRIED odd element afterwards) and the forward
    / backward steps of the round kernel == the XYZZ chain == the oracle's sum, for slices of length
    1/2/3/odd/even/one full 2^R slice, with duplicates (tangent), opposite points, identity bases;
    and a slice of k distinct points costs exactly k - 1 additions."""
    P, Rm = fields.Fp.p, fields.Fp.R
    Ri = pow(Rm, -1, P)
    fl = lambda x: [((x * Rm % P) >> (32 * i)) & 0xFFFFFFFF for i in range(12)]
    fv = lambda l: _v(l) * Ri % P
    rng = random.Random(9)
    if which == "g1":
        G, fn, words = curves.G1, lib.hc_pairs_g1, 24
        enc = lambda p: [0] * 24 if p is None else fl(p[0]) + fl(p[1])
        dec = lambda a: None if not any(a) else (fv(a[:12]), fv(a[12:]))
    else:
        G, fn, words = curves.G2, lib.hc_pairs_g2, 48
        enc = lambda p: [0] * 48 if p is None else fl(p[0][0]) + fl(p[0][1]) + fl(p[1][0]) + fl(p[1][1])
        dec = lambda a: None if not any(a) else ((fv(a[:12]), fv(a[12:24])), (fv(a[24:36]), fv(a[36:])))
    fn.restype = ctypes.c_uint64
    nb = 40
    pts = [G.mul(G.gen, rng.randrange(1, fields.Fr.p)) for _ in range(nb - 1)] + [None]
    bases = (ctypes.c_uint32 * (words * nb))(*sum((enc(p) for p in pts), []))
    R = 6                                            # slices of at most 64 entries
    PAD = 0xFFFFFFFF
    for special in (False, True):
        lens = [1, 2, 3, 4, 5, 7, 8, 9, 17, 31, 33, 63, 64, 64] + [rng.randrange(1, 65) for _ in range(12)]
        entries, tasks, real = [], [], []
        for ln in lens:
            start = len(entries)
            mine = []
            for j in range(ln):
                r = rng.random()
                if special and r < 0.15 and j > 0:
                    e = mine[-1] ^ 0x80000000                 # the negative of the previous point
                elif special and r < 0.3 and j > 0:
                    e = mine[-1]                              # the same point again (tangent)
                elif special:
                    e = rng.randrange(nb) | (rng.randrange(2) << 31)
                else:
                    e = (start + 7 * j) % (nb - 1) if False else rng.randrange(nb - 1) | (rng.randrange(2) << 31)
                mine.append(e)
            if not special:                                   # distinct points in every slice: no special case
                idx = rng.sample(range(nb - 1), min(ln, nb - 1))
                mine = [(idx[j % len(idx)] if ln <= nb - 1 else rng.randrange(nb - 1)) | (rng.randrange(2) << 31)
                        for j in range(ln)]
                if ln > nb - 1:
                    mine = None
            if mine is None:
                continue
            entries += mine + ([PAD] if ln % 2 else [])
            tasks += [start, ln + ln % 2]
            real.append(mine)
        nt = len(real)
        srt = (ctypes.c_uint32 * len(entries))(*entries)
        tk = (ctypes.c_uint32 * (2 * nt))(*tasks)
        for chunk in (1, 5, 1000):
            out_pairs = (ctypes.c_uint32 * (words * nt))()
            out_chain = (ctypes.c_uint32 * (words * nt))()
            adds = fn(bases, srt, len(entries), tk, nt, R, chunk, out_pairs, out_chain)
            assert list(out_pairs) == list(out_chain), (which, special, chunk)
            if not special:
                assert adds == sum(len(m) - 1 for m in real)
            else:
                assert adds <= sum(len(m) - 1 for m in real)
        got = list(out_pairs)
        for t, mine in enumerate(real):
            expect = None
            for e in mine:
                p = pts[e & 0x7FFFFFFF]
                expect = G.add(expect, G.neg(p) if e >> 31 else p)
            assert dec(got[t * words:(t + 1) * words]) == expect, (which, special, t)
