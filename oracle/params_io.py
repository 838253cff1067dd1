"""TEST INFRASTRUCTURE ONLY -- restatement of the reference's Parameters / VerifyingKey / Proof
wire formats:

  Proof::{write,read} ............ src/groth16/mod.rs:42-102   (3 compressed points, 192 B)
  VerifyingKey::{write,read} ..... src/groth16/mod.rs:146-221  (6 uncompressed points, BE u32 |ic|, ic)
  Parameters::{write,read} ....... src/groth16/mod.rs:261-400  (vk, then h, l, a, b_g1, b_g2 each as
                                   BE u32 length + uncompressed points; `checked` selects
                                   from_uncompressed vs from_uncompressed_unchecked; identity rejected)

Point decoding follows the published ZCash BLS12-381 encoding implemented by bls12_381 0.6.0
(third-party): flag bits compression|infinity|sort in the top three bits of byte 0, canonical
big-endian coordinates (G2: c1 before c0).  Known answer from the reference's own test
(groth16/mod.rs:532): the 1-constraint MySillyCircuit serialises to 2136 bytes.
"""
from __future__ import annotations

import struct

from . import curves, fields
from .groth16 import Parameters, VerifyingKey

P = fields.FP_MODULUS
R_ORDER = fields.FR_MODULUS


class InvalidData(Exception):        # io::ErrorKind::InvalidData
    pass


class UnexpectedEof(Exception):      # io::ErrorKind::UnexpectedEof (read_exact)
    pass


def _fp(b):
    v = int.from_bytes(b, "big")
    if v >= P:
        raise InvalidData("non-canonical field element")
    return v


def decode_uncompressed(G, raw, checked):
    """G1Affine::from_uncompressed / from_uncompressed_unchecked (bls12_381 0.6.0)"""
    n = G.coord_bytes
    if len(raw) < 2 * n:
        raise UnexpectedEof()
    comp, inf, sort = raw[0] >> 7 & 1, raw[0] >> 6 & 1, raw[0] >> 5 & 1
    body = bytes([raw[0] & 0x1F]) + bytes(raw[1:2 * n])
    name = "invalid " + G.name
    try:
        if n == 48:
            x, y = _fp(body[:48]), _fp(body[48:96])
            zero = x == 0 and y == 0
        else:
            x = (_fp(body[48:96]), _fp(body[0:48]))
            y = (_fp(body[144:192]), _fp(body[96:144]))
            zero = x == (0, 0) and y == (0, 0)
    except InvalidData:
        raise InvalidData(name)
    if comp or sort:
        raise InvalidData(name)
    if inf:
        if not zero:
            raise InvalidData(name)
        return None
    pt = (x, y)
    if checked and not (G.is_on_curve(pt) and G.mul(pt, R_ORDER) is None):
        raise InvalidData(name)
    return pt


class _Reader:
    def __init__(self, data):
        self.data, self.pos = bytes(data), 0

    def take(self, n):
        if self.pos + n > len(self.data):
            raise UnexpectedEof()
        out = self.data[self.pos:self.pos + n]
        self.pos += n
        return out

    def u32(self):
        return struct.unpack(">I", self.take(4))[0]


def write_vk(vk):
    G1, G2 = curves.G1, curves.G2
    out = (G1.to_uncompressed(vk.alpha_g1) + G1.to_uncompressed(vk.beta_g1) + G2.to_uncompressed(vk.beta_g2)
           + G2.to_uncompressed(vk.gamma_g2) + G1.to_uncompressed(vk.delta_g1) + G2.to_uncompressed(vk.delta_g2))
    out += struct.pack(">I", len(vk.ic)) + b"".join(G1.to_uncompressed(p) for p in vk.ic)
    return out


def write_parameters(params):
    G1, G2 = curves.G1, curves.G2
    out = write_vk(params.vk)
    for G, vec in ((G1, params.h), (G1, params.l), (G1, params.a), (G1, params.b_g1), (G2, params.b_g2)):
        out += struct.pack(">I", len(vec)) + b"".join(G.to_uncompressed(p) for p in vec)
    return out


def _read_point(rd, G, checked, reject_identity):
    pt = decode_uncompressed(G, rd.take(2 * G.coord_bytes), checked)
    if reject_identity and pt is None:
        raise InvalidData("point at infinity")
    return pt


def read_vk(rd):
    G1, G2 = curves.G1, curves.G2
    alpha_g1 = _read_point(rd, G1, True, False)
    beta_g1 = _read_point(rd, G1, True, False)
    beta_g2 = _read_point(rd, G2, True, False)
    gamma_g2 = _read_point(rd, G2, True, False)
    delta_g1 = _read_point(rd, G1, True, False)
    delta_g2 = _read_point(rd, G2, True, False)
    ic = [_read_point(rd, G1, True, True) for _ in range(rd.u32())]
    return VerifyingKey(alpha_g1=alpha_g1, beta_g1=beta_g1, beta_g2=beta_g2, gamma_g2=gamma_g2,
                        delta_g1=delta_g1, delta_g2=delta_g2, ic=ic)


def read_parameters(data, checked):
    G1, G2 = curves.G1, curves.G2
    rd = _Reader(data)
    vk = read_vk(rd)
    vecs = []
    for G in (G1, G1, G1, G1, G2):
        vecs.append([_read_point(rd, G, checked, True) for _ in range(rd.u32())])
    return Parameters(vk, *vecs)


def my_silly_circuit(a, b):
    """groth16/mod.rs:491-518"""
    def synth(cs):
        av = cs.alloc(lambda: a)
        bv = cs.alloc(lambda: b)
        cv = cs.alloc_input(lambda: None if a is None or b is None else a * b % fields.FR_MODULUS)
        cs.enforce([(av, 1)], [(bv, 1)], [(cv, 1)])
    return synth
