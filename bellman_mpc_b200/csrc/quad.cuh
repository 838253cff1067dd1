// Group operations shared by FOUR lanes (a "quad" = lanes 4i .. 4i+3 of a warp), for the
// latency-bound tails of the multiexp: the fold / final kernels of the bucket reduction walk a
// serial chain of ~25 additions and ~20 doublings on a handful of warps, and a chain step costs the
// latency of its 14 (9) dependent-looking field products, 19 (12) us with the call-based product.
// But the products of one XYZZ addition form a dependency graph only 4 levels deep (6, 2, 3, 3
// products wide; doubling: 3 levels, 2, 4, 3 wide).  Here every lane of the quad holds a full copy
// of the operands, runs ONE product of the current level, and the results travel by warp shuffles
// (12 words per Fp value); the cheap additions / subtractions are done redundantly by all four.
// An addition takes 5 product latencies instead of 14, a doubling 3 instead of 9.  Same formulas as
// curve.cuh (add-2008-s, dbl-2008-s-1), same complete handling of the special cases: every lane sees
// the same operands, so the quad branches together and falls back to the one-thread body there.
// Requires all four lanes of the quad to be active and to call with identical operands; blockDim.x
// must be a multiple of 4 and quads must not straddle warps (threadIdx.x & 3 is the role).
#pragma once
#include "curve.cuh"

namespace bmpc {

#if defined(__CUDACC__)
template <class T>
__device__ __forceinline__ T quad_bcast(const T& v, uint32_t src_role) {
    static_assert(sizeof(T) % 4 == 0, "word-sized shuffles");
    T r;
    const uint32_t* in = reinterpret_cast<const uint32_t*>(&v);
    uint32_t* out = reinterpret_cast<uint32_t*>(&r);
    const uint32_t base = threadIdx.x & 28u;               // first lane of this quad
    const uint32_t mask = 0xfu << base;                     // only the quad has to be converged
    const int src = (int)(base | src_role);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 4); i++) out[i] = __shfl_sync(mask, in[i], src);
    return r;
}

template <class F>
__device__ __forceinline__ F quad_sel(uint32_t role, const F& a0, const F& a1, const F& a2, const F& a3) {
    F r = a0;
    if (role == 1) r = a1;
    if (role == 2) r = a2;
    if (role == 3) r = a3;
    return r;
}

// a += b.  The four lanes of the quad must call together with identical operands (the shuffles name
// the quad's lanes only, so different quads of a warp may diverge); a quad whose operands make the
// addition trivial takes the one-thread paths, all four lanes alike.
template <class F>
__device__ __noinline__ void quad_add(XYZZ<F>& a, const XYZZ<F>& b) {
    const uint32_t role = threadIdx.x & 3u;
    const bool b_id = b.is_identity(), a_id = a.is_identity();
    // level 1: U1 = X1 ZZ2 | U2 = X2 ZZ1 | S1 = Y1 ZZZ2 | S2 = Y2 ZZZ1, then ZZ1 ZZ2 | ZZZ1 ZZZ2
    F p = F::mul_cold(quad_sel<F>(role, a.X, b.X, a.Y, b.Y), quad_sel<F>(role, b.ZZ, a.ZZ, b.ZZZ, a.ZZZ));
    const F U1 = quad_bcast(p, 0), U2 = quad_bcast(p, 1), S1 = quad_bcast(p, 2), S2 = quad_bcast(p, 3);
    F t = F::mul_cold(quad_sel<F>(role & 1u, a.ZZ, a.ZZZ, a.ZZ, a.ZZZ), quad_sel<F>(role & 1u, b.ZZ, b.ZZZ, b.ZZ, b.ZZZ));
    const F T1 = quad_bcast(t, 0), T2 = quad_bcast(t, 1);
    const F Pd = U2 - U1, R = S2 - S1;
    // level 2: PP = Pd^2 | RR = R^2
    F q = F::mul_cold(quad_sel<F>(role & 1u, Pd, R, Pd, R), quad_sel<F>(role & 1u, Pd, R, Pd, R));
    const F PP = quad_bcast(q, 0), RR = quad_bcast(q, 1);
    // level 3: PPP = Pd PP | Q = U1 PP | ZZ3 = T1 PP
    F u = F::mul_cold(quad_sel<F>(role, Pd, U1, T1, T1), PP);
    const F PPP = quad_bcast(u, 0), Q = quad_bcast(u, 1), ZZ3 = quad_bcast(u, 2);
    const F X3 = RR - PPP - Q.dbl();
    // level 4: R (Q - X3) | S1 PPP | ZZZ3 = T2 PPP
    F v = F::mul_cold(quad_sel<F>(role, R, S1, T2, T2), quad_sel<F>(role, Q - X3, PPP, PPP, PPP));
    const F Ya = quad_bcast(v, 0), Yb = quad_bcast(v, 1), ZZZ3 = quad_bcast(v, 2);
    // the special cases, decided on values every lane of the quad holds
    if (b_id) return;
    if (a_id) { a = b; return; }
    if (Pd.is_zero()) {
        if (R.is_zero()) a = a.dbl();
        else a = XYZZ<F>::identity();
        return;
    }
    a.X = X3;
    a.Y = Ya - Yb;
    a.ZZ = ZZ3;
    a.ZZZ = ZZZ3;
}

// 2 a
template <class F>
__device__ __noinline__ XYZZ<F> quad_dbl(const XYZZ<F>& a) {
    const uint32_t role = threadIdx.x & 3u;
    const F U = a.Y.dbl();
    // level 1: V = U^2 | xx = X^2
    F p = F::mul_cold(quad_sel<F>(role & 1u, U, a.X, U, a.X), quad_sel<F>(role & 1u, U, a.X, U, a.X));
    const F V = quad_bcast(p, 0), xx = quad_bcast(p, 1);
    const F M = xx.dbl() + xx;
    // level 2: W = U V | S = X V | MM = M^2 | ZZ3 = V ZZ
    F q = F::mul_cold(quad_sel<F>(role, U, a.X, M, V), quad_sel<F>(role, V, V, M, a.ZZ));
    const F W = quad_bcast(q, 0), S = quad_bcast(q, 1), MM = quad_bcast(q, 2), ZZ3 = quad_bcast(q, 3);
    const F X3 = MM - S.dbl();
    // level 3: M (S - X3) | W Y | ZZZ3 = W ZZZ
    F r = F::mul_cold(quad_sel<F>(role, M, W, W, W), quad_sel<F>(role, S - X3, a.Y, a.ZZZ, a.ZZZ));
    const F Ya = quad_bcast(r, 0), Yb = quad_bcast(r, 1), ZZZ3 = quad_bcast(r, 2);
    if (a.is_identity()) return a;
    return XYZZ<F>{X3, Ya - Yb, ZZ3, ZZZ3};
}
#endif

}  // namespace bmpc
