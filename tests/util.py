"""Shared helpers for the parity tests (test infrastructure)."""
import random

import numpy as np

import bellman_mpc_b200 as bm
from oracle import curves, fields

Q = fields.Fr.p


def rand_scalars(n, seed, kind="uniform"):
    rng = random.Random(seed)
    out = []
    for _ in range(n):
        if kind == "uniform":
            out.append(rng.randrange(Q))
        elif kind == "mixed":       # 30 % zero / 30 % one / 40 % uniform
            u = rng.random()
            out.append(0 if u < 0.3 else 1 if u < 0.6 else rng.randrange(Q))
        elif kind == "small":
            out.append(rng.randrange(1 << 16))
        else:
            raise ValueError(kind)
    return out


def known_dlog_bases(worker, group, ks):
    """P_i = k_i * G computed on the GPU (fixed-base path); returns Bases"""
    G = curves.G1 if group == bm.G1 else curves.G2
    return bm.Bases.fixed_base_mul(worker, group, G.to_uncompressed(G.gen), bm.ints_to_limbs(ks))


def expected_from_dlogs(group, ks, scalars, bits=None, start=0):
    """(sum k_{start+rank} * s_i) * G -- exact expectation without a CPU MSM."""
    G = curves.G1 if group == bm.G1 else curves.G2
    acc, k = 0, start
    for i, s in enumerate(scalars):
        if bits is None or bits[i]:
            acc += ks[k] * s
            k += 1
    return G.to_uncompressed(G.mul(G.gen, acc % Q))


def decode(group, data):
    G = curves.G1 if group == bm.G1 else curves.G2
    return G.from_uncompressed(bytes(data))
