// BLS12-381 G1 / G2 group law in extended Jacobian (XYZZ) coordinates, templated on the
// coordinate field (Fp for G1, Fp2 for G2).
//
// Replaces the third-party `G1Projective` / `G2Projective` `add_assign` (mixed and full),
// `double`, `identity`, `is_identity` that the reference calls at
// src/multiexp.rs:39,63,179,185,217,231-232,248 and the scalar multiplications at
// src/groth16/prover.rs:315-343.  The reference's formulas are complete; so are these:
// every entry point handles identity operands, P + P and P + (-P) explicitly, because the
// bucket sort makes equal and opposite points meet (duplicate bases, signed digits).
//
// XYZZ: x = X / ZZ, y = Y / ZZZ with ZZ^3 == ZZZ^2; identity <=> ZZ == 0.
// Formulas: Explicit-Formulas Database, short Weierstrass a = 0, "xyzz" (Sutherland 2008):
// madd-2008-s (8M+2S), add-2008-s (12M+2S), dbl-2008-s-1, mdbl-2008-s-1.
#pragma once
#include "field.cuh"

namespace bmpc {

// Affine point as stored in HBM: Montgomery coordinates, identity encoded as (0, 0)
// (not on either curve since b != 0).  G1: 96 B, G2: 192 B.
template <class F>
struct Affine {
    F x, y;
    BMPC_HD bool is_identity() const { return x.is_zero() && y.is_zero(); }
    BMPC_HD static Affine identity() { return Affine{F::zero(), F::zero()}; }
    BMPC_HD Affine neg() const { return Affine{x, y.neg()}; }
};

template <class F>
struct XYZZ {
    F X, Y, ZZ, ZZZ;

    BMPC_HD static XYZZ identity() { return XYZZ{F::zero(), F::zero(), F::zero(), F::zero()}; }
    BMPC_HD bool is_identity() const { return ZZ.is_zero(); }
    BMPC_HD static XYZZ from_affine(const Affine<F>& p) {
        if (p.is_identity()) return identity();
        return XYZZ{p.x, p.y, F::one(), F::one()};
    }
    BMPC_HD XYZZ neg() const { return XYZZ{X, Y.neg(), ZZ, ZZZ}; }

    // 2 * (affine p), p != identity
    BMPC_COLD static XYZZ dbl_affine(const Affine<F>& p) {
        F U = p.y.dbl();
        F V = U.sqr();
        F W = U * V;
        F S = p.x * V;
        F xx = p.x.sqr();
        F M = xx.dbl() + xx;
        F X3 = M.sqr() - S.dbl();
        F Y3 = M * (S - X3) - W * p.y;
        return XYZZ{X3, Y3, V, W};  // y == 0 gives ZZ == 0 == identity, as it must
    }

    BMPC_COLD XYZZ dbl() const {
        if (is_identity()) return *this;
        F U = Y.dbl();
        F V = U.sqr();
        F W = U * V;
        F S = X * V;
        F xx = X.sqr();
        F M = xx.dbl() + xx;
        F X3 = M.sqr() - S.dbl();
        F Y3 = M * (S - X3) - W * Y;
        return XYZZ{X3, Y3, V * ZZ, W * ZZZ};
    }

    // out-of-line alias for cold callers
    BMPC_COLD void add_affine_cold(const Affine<F>& p) { add_affine(p); }

    // this += affine p (complete); inlined into the bucket-accumulation loop
    BMPC_HD void add_affine(const Affine<F>& p) {
        if (p.is_identity()) return;
        if (is_identity()) {
            X = p.x; Y = p.y; ZZ = F::one(); ZZZ = F::one();
            return;
        }
        F U2 = p.x * ZZ;
        F S2 = p.y * ZZZ;
        F Pd = U2 - X;
        F R = S2 - Y;
        if (Pd.is_zero()) {
            if (R.is_zero()) *this = dbl_affine(p);
            else *this = identity();
            return;
        }
        F PP = Pd.sqr();
        F PPP = Pd * PP;
        F Q = X * PP;
        F X3 = R.sqr() - PPP - Q.dbl();
        Y = R * (Q - X3) - Y * PPP;
        X = X3;
        ZZ = ZZ * PP;
        ZZZ = ZZZ * PPP;
    }

    // this += o (complete), out of line for cold callers
    BMPC_COLD void add(const XYZZ& o) { add_inl(o); }

    // this += o (complete), inlined into the bucket-reduction loop
    BMPC_HD void add_inl(const XYZZ& o) {
        if (o.is_identity()) return;
        if (is_identity()) { *this = o; return; }
        F U1 = X * o.ZZ;
        F U2 = o.X * ZZ;
        F S1 = Y * o.ZZZ;
        F S2 = o.Y * ZZZ;
        F Pd = U2 - U1;
        F R = S2 - S1;
        if (Pd.is_zero()) {
            if (R.is_zero()) *this = dbl();
            else *this = identity();
            return;
        }
        F PP = Pd.sqr();
        F PPP = Pd * PP;
        F Q = U1 * PP;
        F X3 = R.sqr() - PPP - Q.dbl();
        Y = R * (Q - X3) - S1 * PPP;
        X = X3;
        ZZ = ZZ * o.ZZ * PP;
        ZZZ = ZZZ * o.ZZZ * PPP;
    }

    // canonical affine coordinates (one field inversion)
    BMPC_COLD Affine<F> to_affine() const {
        if (is_identity()) return Affine<F>::identity();
        F i = F::mul_cold(ZZ, ZZZ).inv();     // 1 / (ZZ * ZZZ)
        F izz = F::mul_cold(i, ZZZ);          // 1 / ZZ
        F izzz = F::mul_cold(i, ZZ);          // 1 / ZZZ
        return Affine<F>{F::mul_cold(X, izz), F::mul_cold(Y, izzz)};
    }

    // this * k for a little-endian multi-word scalar (double-and-add, vartime)
    BMPC_COLD XYZZ mul(const uint32_t* k, int words) const {
        XYZZ r = identity();
        for (int i = words - 1; i >= 0; i--) {
            for (int b = 31; b >= 0; b--) {
                r = r.dbl();
                if ((k[i] >> b) & 1) r.add(*this);
            }
        }
        return r;
    }
};

typedef Affine<Fp> G1Affine;
typedef Affine<Fp2> G2Affine;
typedef XYZZ<Fp> G1XYZZ;
typedef XYZZ<Fp2> G2XYZZ;

}  // namespace bmpc
