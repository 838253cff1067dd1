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

    // Every group operation exists in two flavours sharing one body (template on COMPACT):
    //  * inlined field products -- for the hot bucket-accumulation loop, where the ~7000
    //    instruction straight-line body streams well because all warps run it in step;
    //  * COMPACT: every field product is a call of the single out-of-line F::mul_cold.  ncu showed
    //    the cold kernels (bucket reduction, heavy-bucket combine, final fold, proof tail) starved
    //    for instructions (stall "no_instruction" 7.7 per issue): with a few divergent warps per SM
    //    a 130 KB addition body thrashes the 32 KB instruction cache.  The compact bodies are ~1 K
    //    instructions plus one 10 KB product routine.
    template <bool COMPACT>
    BMPC_HD static F M(const F& a, const F& b) {
        if (COMPACT) return F::mul_cold(a, b);
        return a * b;
    }
    // square: for Fp2 the inlined path uses complex squaring (2 Fp products instead of 3)
    template <bool COMPACT>
    BMPC_HD static F SQ(const F& a) {
        if (COMPACT) return F::mul_cold(a, a);
        return a.sqr();
    }

    // 2 * (affine p), p != identity
    template <bool COMPACT>
    BMPC_HD static XYZZ dbl_affine_impl(const Affine<F>& p) {
        F U = p.y.dbl();
        F V = M<COMPACT>(U, U);
        F W = M<COMPACT>(U, V);
        F S = M<COMPACT>(p.x, V);
        F xx = M<COMPACT>(p.x, p.x);
        F Mm = xx.dbl() + xx;
        F X3 = M<COMPACT>(Mm, Mm) - S.dbl();
        F Y3 = M<COMPACT>(Mm, S - X3) - M<COMPACT>(W, p.y);
        return XYZZ{X3, Y3, V, W};  // y == 0 gives ZZ == 0 == identity, as it must
    }
    BMPC_COLD static XYZZ dbl_affine(const Affine<F>& p) { return dbl_affine_impl<true>(p); }

    BMPC_COLD XYZZ dbl() const {
        if (is_identity()) return *this;
        F U = Y.dbl();
        F V = M<true>(U, U);
        F W = M<true>(U, V);
        F S = M<true>(X, V);
        F xx = M<true>(X, X);
        F Mm = xx.dbl() + xx;
        F X3 = M<true>(Mm, Mm) - S.dbl();
        F Y3 = M<true>(Mm, S - X3) - M<true>(W, Y);
        return XYZZ{X3, Y3, M<true>(V, ZZ), M<true>(W, ZZZ)};
    }

    // this += affine p (complete)
    template <bool COMPACT>
    BMPC_HD void add_affine_impl(const Affine<F>& p) {
        if (p.is_identity()) return;
        if (is_identity()) {
            X = p.x; Y = p.y; ZZ = F::one(); ZZZ = F::one();
            return;
        }
        F U2 = M<COMPACT>(p.x, ZZ);
        F S2 = M<COMPACT>(p.y, ZZZ);
        F Pd = U2 - X;
        F R = S2 - Y;
        if (Pd.is_zero()) {
            if (R.is_zero()) *this = dbl_affine(p);
            else *this = identity();
            return;
        }
        F PP = SQ<COMPACT>(Pd);
        F PPP = M<COMPACT>(Pd, PP);
        F Q = M<COMPACT>(X, PP);
        F X3 = SQ<COMPACT>(R) - PPP - Q.dbl();
        Y = M<COMPACT>(R, Q - X3) - M<COMPACT>(Y, PPP);
        X = X3;
        ZZ = M<COMPACT>(ZZ, PP);
        ZZZ = M<COMPACT>(ZZZ, PPP);
    }
    // inlined into the bucket-accumulation loop
    BMPC_HD void add_affine(const Affine<F>& p) { add_affine_impl<false>(p); }
    // out-of-line, compact: cold callers
    BMPC_COLD void add_affine_cold(const Affine<F>& p) { add_affine_impl<true>(p); }

    // this += o (complete)
    template <bool COMPACT>
    BMPC_HD void add_impl(const XYZZ& o) {
        if (o.is_identity()) return;
        if (is_identity()) { *this = o; return; }
        F U1 = M<COMPACT>(X, o.ZZ);
        F U2 = M<COMPACT>(o.X, ZZ);
        F S1 = M<COMPACT>(Y, o.ZZZ);
        F S2 = M<COMPACT>(o.Y, ZZZ);
        F Pd = U2 - U1;
        F R = S2 - S1;
        if (Pd.is_zero()) {
            if (R.is_zero()) *this = dbl();
            else *this = identity();
            return;
        }
        F PP = M<COMPACT>(Pd, Pd);
        F PPP = M<COMPACT>(Pd, PP);
        F Q = M<COMPACT>(U1, PP);
        F X3 = M<COMPACT>(R, R) - PPP - Q.dbl();
        Y = M<COMPACT>(R, Q - X3) - M<COMPACT>(S1, PPP);
        X = X3;
        ZZ = M<COMPACT>(M<COMPACT>(ZZ, o.ZZ), PP);
        ZZZ = M<COMPACT>(M<COMPACT>(ZZZ, o.ZZZ), PPP);
    }
    BMPC_COLD void add(const XYZZ& o) { add_impl<true>(o); }      // compact, out of line
    BMPC_HD void add_inl(const XYZZ& o) { add_impl<false>(o); }    // inlined products

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
