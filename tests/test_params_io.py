"""Parameters / VerifyingKey wire format (groth16/mod.rs:146-221,261-400): the oracle must
reproduce the reference's known size (2136 B, mod.rs:532), round-trip, and reject what the
reference rejects; the GPU ingestion (bmpc_params_read) must agree with it byte for byte."""
import struct

import pytest

from oracle import curves, fields
from oracle import groth16 as og
from oracle import params_io as pio


def _params():
    return og.generate_random_parameters(og.BLS12, pio.my_silly_circuit(None, None))


def _same(a, b):
    return (a.h, a.l, a.a, a.b_g1, a.b_g2, a.vk.ic, a.vk.alpha_g1, a.vk.beta_g2, a.vk.gamma_g2, a.vk.delta_g2) == \
           (b.h, b.l, b.a, b.b_g1, b.b_g2, b.vk.ic, b.vk.alpha_g1, b.vk.beta_g2, b.vk.gamma_g2, b.vk.delta_g2)


def test_known_size_and_roundtrip():
    params = _params()
    blob = pio.write_parameters(params)
    assert len(blob) == 2136                                       # mod.rs:532
    assert (len(params.vk.ic), len(params.h), len(params.l), len(params.a), len(params.b_g1), len(params.b_g2)) == (2, 3, 2, 3, 1, 1)
    assert _same(pio.read_parameters(blob, True), params)          # :534-535
    assert _same(pio.read_parameters(blob, False), params)         # :537-538
    assert pio.write_parameters(pio.read_parameters(blob, False)) == blob


def corruptions(blob):
    """(name, corrupted blob, error for checked, error for unchecked)"""
    h0 = 864 + 4 + 2 * 96 + 4                     # first h point
    out = []
    b = bytearray(blob); b[h0] |= 0x80
    out.append(("compression flag on an uncompressed point", bytes(b), pio.InvalidData, pio.InvalidData))
    b = bytearray(blob); b[h0:h0 + 96] = bytes([0x40]) + bytes(95)
    out.append(("identity in h", bytes(b), pio.InvalidData, pio.InvalidData))
    b = bytearray(blob); b[h0 + 95] ^= 1
    out.append(("h point off the curve", bytes(b), pio.InvalidData, None))
    b = bytearray(blob); b[h0:h0 + 48] = (fields.FP_MODULUS).to_bytes(48, "big")
    out.append(("non-canonical x", bytes(b), pio.InvalidData, pio.InvalidData))
    out.append(("truncated", blob[:-10], pio.UnexpectedEof, pio.UnexpectedEof))
    b = bytearray(blob); b[864:868] = struct.pack(">I", 3)
    out.append(("ic length too long: the next bytes are not a point", bytes(b), (pio.InvalidData, pio.UnexpectedEof), (pio.InvalidData, pio.UnexpectedEof)))
    # a point on the curve but outside the prime-order subgroup
    x = 1
    while True:
        y2 = (x ** 3 + 4) % fields.FP_MODULUS
        y = pow(y2, (fields.FP_MODULUS + 1) // 4, fields.FP_MODULUS)
        if y * y % fields.FP_MODULUS == y2 and curves.G1.mul((x, y), fields.FR_MODULUS) is not None:
            break
        x += 1
    b = bytearray(blob); b[h0:h0 + 96] = x.to_bytes(48, "big") + y.to_bytes(48, "big")
    out.append(("on curve, not in the subgroup", bytes(b), pio.InvalidData, None))
    return out


def test_oracle_rejections():
    blob = pio.write_parameters(_params())
    for name, bad, err_checked, err_unchecked in corruptions(blob):
        for checked, err in ((True, err_checked), (False, err_unchecked)):
            if err is None:
                pio.read_parameters(bad, checked)
            else:
                with pytest.raises(err):
                    pio.read_parameters(bad, checked)


@pytest.mark.gpu
def test_gpu_params_read_write(worker):
    import bellman_mpc_b200 as bm
    params = _params()
    blob = pio.write_parameters(params)
    for checked in (True, False):
        gp = bm.Parameters.read(worker, blob, checked)
        assert len(gp.h) == 3 and len(gp.l) == 2 and len(gp.a) == 3 and len(gp.b_g1) == 1 and len(gp.b_g2) == 1
        assert gp.write() == blob
        # and the resident CRS proves: same bytes as the oracle
        pr = og.synthesize_for_proving(og.BLS12, pio.my_silly_circuit(3, 5))
        from test_gpu_prove import to_gpu_assignment
        proof = bm.create_random_proof(to_gpu_assignment(pr), gp)
        assert proof == og.expected_proof(og.BLS12, params, pr, 27134, 17146).to_bytes(og.BLS12)
        gp.free()
    for name, bad, err_checked, err_unchecked in corruptions(blob):
        for checked, err in ((True, err_checked), (False, err_unchecked)):
            if err is None:
                bm.Parameters.read(worker, bad, checked).free()
                continue
            with pytest.raises((bm.InvalidData, bm.UnexpectedEof)) as ei:
                bm.Parameters.read(worker, bad, checked)
            errs = err if isinstance(err, tuple) else (err,)
            want = tuple(bm.InvalidData if e is pio.InvalidData else bm.UnexpectedEof for e in errs)
            assert isinstance(ei.value, want), name
