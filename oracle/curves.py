"""TEST INFRASTRUCTURE ONLY -- big-integer oracle for the groups on the hot path.

Restates the published BLS12-381 G1/G2 group law (third-party crate bls12_381 0.6.0,
Cargo.lock:96-99; the reference calls it through `add_assign`/`double`/`identity`/
`is_identity` at src/multiexp.rs:39,63,179,185,231-232,248 and through scalar
multiplication / `to_affine` at src/groth16/prover.rs:315-349) and the ZCash point
encodings used by `Proof`/`Parameters` (src/groth16/mod.rs:42-48,261-290).

Also provides the reference's DummyEngine "group" (src/groth16/tests/dummy_engine.rs:
331-365: G1 = G2 = Gt = Fr = Z/64513, generator 1, scalar mul = field mul) so the
same generic multiexp / create_proof restatement can be pinned to the integer golden
vectors in src/groth16/tests/mod.rs.

Points are affine: None = identity, G1 = (x, y), G2 = ((x0, x1), (y0, y1)).
"""
from __future__ import annotations

from .fields import Fp, Fr, FP_MODULUS, DummyFr

P = FP_MODULUS


# ----------------------------------------------------------------------------- Fp2
def fp2_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def fp2_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def fp2_neg(a):
    return ((-a[0]) % P, (-a[1]) % P)


def fp2_mul(a, b):
    # (a0 + a1 u)(b0 + b1 u), u^2 = -1
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def fp2_sqr(a):
    return fp2_mul(a, a)


def fp2_inv(a):
    n = pow((a[0] * a[0] + a[1] * a[1]) % P, -1, P)
    return (a[0] * n % P, (-a[1]) * n % P)


def fp_sqrt(a):
    """p = 3 mod 4: a^((p+1)/4), None when a is not a square"""
    r = pow(a, (P + 1) // 4, P)
    return r if r * r % P == a % P else None


def fp2_sqrt(a):
    """Square root in Fp[u]/(u^2+1) for p = 3 mod 4 (Adj & Rodriguez-Henriquez, alg. 9), None for a
    non-residue."""
    if a == (0, 0):
        return (0, 0)

    def fpow(x, e):
        r = (1, 0)
        while e:
            if e & 1:
                r = fp2_mul(r, x)
            x = fp2_mul(x, x)
            e >>= 1
        return r
    a1 = fpow(a, (P - 3) // 4)
    alpha = fp2_mul(a1, fp2_mul(a1, a))
    x0 = fp2_mul(a1, a)
    if alpha == (P - 1, 0):
        r = fp2_mul((0, 1), x0)
    else:
        r = fp2_mul(fpow(fp2_add((1, 0), alpha), (P - 1) // 2), x0)
    return r if fp2_mul(r, r) == (a[0] % P, a[1] % P) else None


def fp2_scalar(a, k):
    return (a[0] * k % P, a[1] * k % P)


# ------------------------------------------------------------- generic short-Weierstrass
class _FieldOps:
    """Tiny adapter so the same Jacobian formulas serve Fp and Fp2."""

    def __init__(self, add, sub, mul, neg, inv, zero, one, is_zero):
        self.add, self.sub, self.mul, self.neg, self.inv = add, sub, mul, neg, inv
        self.zero, self.one, self.is_zero = zero, one, is_zero


_FP = _FieldOps(lambda a, b: (a + b) % P, lambda a, b: (a - b) % P,
                lambda a, b: a * b % P, lambda a: (-a) % P,
                lambda a: pow(a, -1, P), 0, 1, lambda a: a % P == 0)
_FP2 = _FieldOps(fp2_add, fp2_sub, fp2_mul, fp2_neg, fp2_inv, (0, 0), (1, 0),
                 lambda a: a[0] % P == 0 and a[1] % P == 0)


class CurveGroup:
    """y^2 = x^3 + b over F.  Affine in / affine out; Jacobian inside `mul`/`sum`."""

    def __init__(self, name, F, b, gen, coord_bytes):
        self.name, self.F, self.b, self.gen = name, F, b, gen
        self.coord_bytes = coord_bytes
        self.scalar_field = Fr

    # -- predicates ----------------------------------------------------------
    def identity(self):
        return None

    def is_identity(self, p):
        return p is None

    def is_on_curve(self, p):
        if p is None:
            return True
        F = self.F
        x, y = p
        return F.sub(F.mul(y, y), F.add(F.mul(F.mul(x, x), x), self.b)) == F.zero

    def eq(self, a, b):
        return a == b

    # -- affine group law (complete: handles identity, P+P, P+(-P)) ------------
    def neg(self, p):
        if p is None:
            return None
        return (p[0], self.F.neg(p[1]))

    def add(self, p, q):
        F = self.F
        if p is None:
            return q
        if q is None:
            return p
        x1, y1 = p
        x2, y2 = q
        if x1 == x2:
            if y1 == y2 and not F.is_zero(y1):
                return self.double(p)
            return None
        lam = F.mul(F.sub(y2, y1), F.inv(F.sub(x2, x1)))
        x3 = F.sub(F.sub(F.mul(lam, lam), x1), x2)
        y3 = F.sub(F.mul(lam, F.sub(x1, x3)), y1)
        return (x3, y3)

    def sub(self, p, q):
        return self.add(p, self.neg(q))

    def double(self, p):
        F = self.F
        if p is None:
            return None
        x1, y1 = p
        if F.is_zero(y1):
            return None
        xx = F.mul(x1, x1)
        lam = F.mul(F.add(F.add(xx, xx), xx), F.inv(F.add(y1, y1)))
        x3 = F.sub(F.sub(F.mul(lam, lam), x1), x1)
        y3 = F.sub(F.mul(lam, F.sub(x1, x3)), y1)
        return (x3, y3)

    # -- Jacobian helpers (speed only) ---------------------------------------
    def _jdbl(self, X, Y, Z):
        F = self.F
        if F.is_zero(Z) or F.is_zero(Y):
            return (F.one, F.one, F.zero)
        A = F.mul(X, X)
        B = F.mul(Y, Y)
        C = F.mul(B, B)
        t = F.add(X, B)
        D = F.sub(F.sub(F.mul(t, t), A), C)
        D = F.add(D, D)
        E = F.add(F.add(A, A), A)
        Fq = F.mul(E, E)
        X3 = F.sub(Fq, F.add(D, D))
        C8 = F.add(C, C)
        C8 = F.add(C8, C8)
        C8 = F.add(C8, C8)
        Y3 = F.sub(F.mul(E, F.sub(D, X3)), C8)
        Z3 = F.mul(Y, Z)
        Z3 = F.add(Z3, Z3)
        return (X3, Y3, Z3)

    def _jadd_affine(self, X1, Y1, Z1, q):
        F = self.F
        if q is None:
            return (X1, Y1, Z1)
        x2, y2 = q
        if F.is_zero(Z1):
            return (x2, y2, F.one)
        Z1Z1 = F.mul(Z1, Z1)
        U2 = F.mul(x2, Z1Z1)
        S2 = F.mul(F.mul(y2, Z1), Z1Z1)
        if U2 == X1:
            if S2 == Y1:
                return self._jdbl(X1, Y1, Z1)
            return (F.one, F.one, F.zero)
        H = F.sub(U2, X1)
        HH = F.mul(H, H)
        HHH = F.mul(H, HH)
        r = F.sub(S2, Y1)
        V = F.mul(X1, HH)
        X3 = F.sub(F.sub(F.mul(r, r), HHH), F.add(V, V))
        Y3 = F.sub(F.mul(r, F.sub(V, X3)), F.mul(Y1, HHH))
        Z3 = F.mul(Z1, H)
        return (X3, Y3, Z3)

    def _to_affine(self, X, Y, Z):
        F = self.F
        if F.is_zero(Z):
            return None
        zi = F.inv(Z)
        zi2 = F.mul(zi, zi)
        return (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))

    def mul(self, p, k):
        """Scalar multiplication p * k (k any non-negative int; reduced mod r is the
        caller's business -- the group has prime order r so k and k mod r agree)."""
        F = self.F
        if p is None or k == 0:
            return None
        acc = (F.one, F.one, F.zero)
        for bit in bin(k)[2:]:
            acc = self._jdbl(*acc)
            if bit == "1":
                acc = self._jadd_affine(*acc, p)
        return self._to_affine(*acc)

    def sum(self, pts):
        F = self.F
        acc = (F.one, F.one, F.zero)
        for q in pts:
            acc = self._jadd_affine(*acc, q)
        return self._to_affine(*acc)

    # -- ZCash encodings (groth16/mod.rs:42-48,146-159,261-290) ------------------
    def _coord_to_bytes(self, c):
        if isinstance(c, tuple):   # Fp2: c1 first, then c0 (same order gt_bytes.rs:41-59 prints)
            return c[1].to_bytes(48, "big") + c[0].to_bytes(48, "big")
        return c.to_bytes(48, "big")

    def _coord_from_bytes(self, b):
        if self.coord_bytes == 96:
            return (int.from_bytes(b[48:96], "big"), int.from_bytes(b[0:48], "big"))
        return int.from_bytes(b, "big")

    def _lex_largest(self, y):
        """bls12_381 `lexicographically_largest`: y > -y as integers (Fp2: compare c1 first)."""
        if isinstance(y, tuple):
            if y[1] != 0:
                return y[1] > (P - 1) // 2
            return y[0] > (P - 1) // 2
        return y > (P - 1) // 2

    def to_uncompressed(self, p):
        n = self.coord_bytes
        if p is None:
            out = bytearray(2 * n)
            out[0] |= 0x40
            return bytes(out)
        return self._coord_to_bytes(p[0]) + self._coord_to_bytes(p[1])

    def from_uncompressed(self, b):
        n = self.coord_bytes
        assert len(b) == 2 * n
        if b[0] & 0x40:
            return None
        b = bytes([b[0] & 0x1F]) + bytes(b[1:])
        return (self._coord_from_bytes(b[:n]), self._coord_from_bytes(b[n:]))

    def to_compressed(self, p):
        n = self.coord_bytes
        if p is None:
            out = bytearray(n)
            out[0] |= 0xC0
            return bytes(out)
        out = bytearray(self._coord_to_bytes(p[0]))
        out[0] |= 0x80
        if self._lex_largest(p[1]):
            out[0] |= 0x20
        return bytes(out)


    def from_compressed(self, b):
        """bls12_381 `from_compressed` (the GroupEncoding::from_bytes that Proof::read calls,
        groth16/mod.rs:50-103): flag bits, canonical coordinates, on-curve by construction of y,
        sign bit, and the prime-order subgroup check.  None = the CtOption is none."""
        n = self.coord_bytes
        if len(b) != n or not (b[0] & 0x80):
            return None
        inf, sign = bool(b[0] & 0x40), bool(b[0] & 0x20)
        body = bytes([b[0] & 0x1F]) + bytes(b[1:])
        if n == 96:
            c1, c0 = int.from_bytes(body[:48], "big"), int.from_bytes(body[48:], "big")
            if c0 >= P or c1 >= P:
                return None
            x = (c0, c1)
        else:
            x = int.from_bytes(body, "big")
            if x >= P:
                return None
        if inf:
            zero = x == ((0, 0) if n == 96 else 0)
            return "identity" if (zero and not sign) else None
        F = self.F
        rhs = F.add(F.mul(F.mul(x, x), x), self.b)
        y = fp2_sqrt(rhs) if n == 96 else fp_sqrt(rhs)
        if y is None:
            return None
        if self._lex_largest(y) != sign:
            y = F.neg(y)
        pt = (x, y)
        if self.mul(pt, Fr.p) is not None:
            return None
        return pt


G1_GEN = (
    0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
    0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1,
)
G2_GEN = (
    (0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
     0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E),
    (0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
     0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE),
)

G1 = CurveGroup("G1", _FP, 4, G1_GEN, 48)
G2 = CurveGroup("G2", _FP2, (4, 4), G2_GEN, 96)


class DummyGroup:
    """DummyEngine's G1 = G2 = Fr (dummy_engine.rs:331-365): additive group of Z/64513,
    identity 0, generator 1, `p * k` = field multiplication."""

    name = "Dummy"
    gen = 1
    scalar_field = DummyFr

    def identity(self):
        return 0

    def is_identity(self, p):
        return p == 0

    def add(self, a, b):
        return (a + b) % 64513

    def sub(self, a, b):
        return (a - b) % 64513

    def neg(self, a):
        return (-a) % 64513

    def double(self, a):
        return 2 * a % 64513

    def mul(self, a, k):
        return a * k % 64513

    def sum(self, pts):
        return sum(pts) % 64513

    def eq(self, a, b):
        return a == b


Dummy = DummyGroup()


def self_check():
    assert G1.is_on_curve(G1_GEN) and G2.is_on_curve(G2_GEN)
    r = Fr.p
    assert G1.mul(G1_GEN, r) is None and G2.mul(G2_GEN, r) is None
    # group-law consistency on the exceptional cases
    two = G1.double(G1_GEN)
    assert G1.add(G1_GEN, G1_GEN) == two and G1.mul(G1_GEN, 2) == two
    assert G1.add(two, G1.neg(two)) is None
    assert G1.add(None, two) == two and G1.add(two, None) == two
    assert G2.add(G2_GEN, G2_GEN) == G2.mul(G2_GEN, 2)
    # encodings: known compressed generator prefix (ZCash spec test value)
    assert G1.to_compressed(G1_GEN).hex().startswith("97f1d3a73197d794")
    assert G1.from_uncompressed(G1.to_uncompressed(two)) == two
    assert G2.from_uncompressed(G2.to_uncompressed(G2_GEN)) == G2_GEN
    for G in (G1, G2):
        for k in (1, 2, 3, 0xDEADBEEF):
            pt = G.mul(G.gen, k)
            assert G.from_compressed(G.to_compressed(pt)) == pt
        assert G.from_compressed(G.to_compressed(None)) == "identity"
    return True
