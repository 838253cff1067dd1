"""TEST INFRASTRUCTURE ONLY -- big-integer oracle for the prime fields on the hot path.

Nothing under oracle/ is product code: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference leg may import it.

The arithmetic of the reference lives in un-vendored third-party crates
(bls12_381 0.6.0, ff 0.11.0 -- Cargo.lock:96-99,291-294).  This file restates
the *published* parameters of those fields and re-derives every constant from
first principles in `self_check()`; the reference's own in-tree copies
(src/gt_bytes.rs:20-30 Fp modulus limbs + INV, src/groth16/tests/dummy_engine.rs:15,
292-316 dummy field) are asserted there too.

Elements are plain Python ints in [0, p).
"""
from __future__ import annotations


class PrimeField:
    """A prime field with the ff::PrimeField constants the reference uses
    (domain.rs:57-77: S, root_of_unity, multiplicative_generator)."""

    def __init__(self, name, modulus, generator, s, root_of_unity, num_bits, limbs64):
        self.name = name
        self.p = modulus
        self.generator = generator          # PrimeField::multiplicative_generator()
        self.S = s                          # PrimeField::S (2-adicity)
        self.root_of_unity = root_of_unity  # PrimeField::root_of_unity()
        self.NUM_BITS = num_bits            # PrimeField::NUM_BITS
        self.limbs64 = limbs64              # Montgomery limb count (64-bit)
        self.R = (1 << (64 * limbs64)) % modulus
        self.R2 = self.R * self.R % modulus
        self.INV64 = (-pow(modulus, -1, 1 << 64)) % (1 << 64)
        self.INV32 = self.INV64 & 0xFFFFFFFF

    # -- arithmetic -------------------------------------------------------
    def add(self, a, b):
        return (a + b) % self.p

    def sub(self, a, b):
        return (a - b) % self.p

    def mul(self, a, b):
        return a * b % self.p

    def neg(self, a):
        return (-a) % self.p

    def square(self, a):
        return a * a % self.p

    def inv(self, a):
        if a % self.p == 0:
            raise ZeroDivisionError("inverse of zero")
        return pow(a, -1, self.p)

    def pow(self, a, e):
        return pow(a, e, self.p)

    # -- encodings --------------------------------------------------------
    def to_mont(self, a):
        return a * self.R % self.p

    def from_mont(self, a):
        return a * pow(self.R, -1, self.p) % self.p

    def to_limbs64(self, a):
        return [(a >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(self.limbs64)]

    def from_limbs64(self, limbs):
        return sum(int(l) << (64 * i) for i, l in enumerate(limbs))


# --- BLS12-381 scalar field Fr (bls12_381::Scalar) -----------------------------
FR_MODULUS = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
FR_ROOT_OF_UNITY = pow(7, (FR_MODULUS - 1) >> 32, FR_MODULUS)
Fr = PrimeField("bls12_381::Scalar", FR_MODULUS, 7, 32, FR_ROOT_OF_UNITY, 255, 4)

# --- BLS12-381 base field Fp ----------------------------------------------------
FP_MODULUS = int(
    "1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f624"
    "1eabfffeb153ffffb9feffffffffaaab", 16)
Fp = PrimeField("bls12_381::Fp", FP_MODULUS, 2, 1, FP_MODULUS - 1, 381, 6)

# --- the reference's DummyEngine field (groth16/tests/dummy_engine.rs:15,292-316)
DummyFr = PrimeField("DummyEngine::Fr", 64513, 5, 10, 57751, 16, 1)


def self_check():
    """Re-derive the constants; cross-check with the reference's in-tree copies."""
    # SURVEY Appendix A / bls12_381 published values
    assert FR_MODULUS.bit_length() == 255 and FP_MODULUS.bit_length() == 381
    assert (FR_MODULUS - 1) % (1 << 32) == 0 and ((FR_MODULUS - 1) >> 32) % 2 == 1
    assert FR_ROOT_OF_UNITY == 0x16A2A19EDFE81F20D09B681922C813B4B63683508C2280B93829971F439F0D2B
    assert pow(FR_ROOT_OF_UNITY, 1 << 31, FR_MODULUS) == FR_MODULUS - 1
    assert Fr.INV64 == 0xFFFFFFFEFFFFFFFF
    assert Fr.to_limbs64(Fr.R) == [0x00000001FFFFFFFE, 0x5884B7FA00034802,
                                   0x998C4FEFECBC4FF5, 0x1824B159ACC5056F]
    # src/gt_bytes.rs:20-30 (reference's own copy of the Fp modulus and INV)
    assert Fp.to_limbs64(FP_MODULUS) == [
        0xB9FEFFFFFFFFAAAB, 0x1EABFFFEB153FFFF, 0x6730D2A0F6B0F624,
        0x64774B84F38512BF, 0x4B1BA7B6434BACD7, 0x1A0111EA397FE69A]
    assert Fp.INV64 == 0x89F3FFFCFFFCFFFD
    # dummy_engine.rs:15,292-316 and tests/mod.rs:334-342
    assert pow(5, 63, 64513) == 57751
    assert pow(57751, 1 << 10, 64513) == 1 and pow(57751, 1 << 9, 64513) != 1
    assert pow(57751, 1 << 7, 64513) == 20201
    return True
