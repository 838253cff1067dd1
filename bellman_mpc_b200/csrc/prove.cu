// Tail of create_proof (reference: src/groth16/prover.rs:309-349): assemble A, B, C from the
// eight multiexp results and the verifying-key points, normalise, and write the 192-byte
// compressed proof (Proof::write, src/groth16/mod.rs:42-48).  Seven scalar multiplications and three
// inversions, kept on the device so that no group arithmetic runs on the host.  Every chain of
// group operations here is pure latency, so each runs on a QUAD of lanes (quad.cuh: an addition in 5
// product latencies instead of 14, a doubling in 3 instead of 9): with the fork's fixed 15-bit r, s
// the G2 product delta_g2 * s alone took 1.1 ms on one thread; a host passing full-size r, s would
// have waited ~16 ms for it.
#include "encode.cuh"
#include "internal.h"
#include "quad.cuh"

namespace bmpc {

// canonical (non-Montgomery) comparison a > b
__device__ __forceinline__ bool fp_canon_gt(const Fp& a, const Fp& b) {
    for (int j = 11; j >= 0; j--) {
        if (a.l[j] > b.l[j]) return true;
        if (a.l[j] < b.l[j]) return false;
    }
    return false;
}
// bls12_381 `lexicographically_largest`: y > -y
__device__ __forceinline__ bool lex_largest(const Fp& y) {
    return fp_canon_gt(y.from_mont(), y.neg().from_mont());
}
__device__ __forceinline__ bool lex_largest(const Fp2& y) {
    if (!y.c1.is_zero()) return lex_largest(y.c1);
    return lex_largest(y.c0);
}
// ZCash compressed encoding: x big-endian, flags 0x80 compressed | 0x40 infinity | 0x20 y-sign
template <class F>
__device__ __forceinline__ void encode_compressed(const Affine<F>& p, uint8_t* out) {
    const int CB = sizeof(F);
    if (p.is_identity()) {
        for (int j = 0; j < CB; j++) out[j] = 0;
        out[0] = 0xc0;
        return;
    }
    put_coord_be(p.x, out);
    out[0] |= 0x80;
    if (lex_largest(p.y)) out[0] |= 0x20;
}

// Stage 1: the seven scalar multiplications of prover.rs:315-343, one thread block each, in
// parallel (a single thread doing them back to back took 7.5 ms; with full-size r, s it would be
// ~30 ms).  tmp1: 6 G1 results, tmp2: 1 G2 result.
//   0 delta1*r   1 delta1*rs   2 alpha*s   3 beta1*r   4 a_answer*s   5 b1_answer*r   | G2: delta2*s
// base * k (little-endian words), double-and-add from the top set bit down, on one quad
template <class F>
__device__ __forceinline__ XYZZ<F> quad_scalar_mul(const XYZZ<F>& base, const uint32_t* k, int words) {
    int top = words * 32 - 1;
    while (top >= 0 && !((k[top >> 5] >> (top & 31)) & 1u)) top--;
    XYZZ<F> r = XYZZ<F>::identity();
    for (int b = top; b >= 0; b--) {
        r = quad_dbl<F>(r);
        if ((k[b >> 5] >> (b & 31)) & 1u) quad_add<F>(r, base);
    }
    return r;
}

__global__ void prove_tail_mul_kernel(ProveTailArgs A, G1XYZZ* tmp1, G2XYZZ* tmp2) {
    if (threadIdx.x >= 4) return;          // one quad per block
    Fr r = A.rs[0], s = A.rs[1];
    Fr rc = r.from_mont(), sc = s.from_mont(), rsc = (r * s).from_mont();   // :322-323
    const uint32_t job = blockIdx.x;
    if (job == 6) {
        G2XYZZ v = quad_scalar_mul<Fp2>(G2XYZZ::from_affine(A.vk_g2[1]), sc.l, 8);   // delta_g2 * s :317
        if (threadIdx.x == 0) tmp2[0] = v;
        return;
    }
    G1XYZZ base;
    const uint32_t* k;
    switch (job) {
        case 0: base = G1XYZZ::from_affine(A.vk_g1[2]); k = rc.l; break;       // delta_g1 * r  :315
        case 1: base = G1XYZZ::from_affine(A.vk_g1[2]); k = rsc.l; break;      // delta_g1 * rs :325
        case 2: base = G1XYZZ::from_affine(A.vk_g1[0]); k = sc.l; break;       // alpha_g1 * s  :326
        case 3: base = G1XYZZ::from_affine(A.vk_g1[1]); k = rc.l; break;       // beta_g1 * r   :327
        case 4: base = *A.a_inputs; quad_add<Fp>(base, *A.a_aux); k = sc.l; break;       // a_answer * s  :331-332
        default: base = *A.b1_inputs; quad_add<Fp>(base, *A.b1_aux); k = rc.l; break;    // b1_answer * r :341-342
    }
    G1XYZZ v = quad_scalar_mul<Fp>(base, k, 8);
    if (threadIdx.x == 0) tmp1[job] = v;
}

// Stage 2: assemble A, B, C (one block = one quad each), normalise and compress.
__global__ void prove_tail_kernel(ProveTailArgs A, const G1XYZZ* tmp1, const G2XYZZ* tmp2) {
    if (threadIdx.x >= 4) return;
    const bool lead = threadIdx.x == 0;
    if (blockIdx.x == 0) {            // A = delta*r + alpha + (a_inputs + a_aux)      :315-316,328-330
        G1XYZZ g_a = tmp1[0];
        quad_add<Fp>(g_a, G1XYZZ::from_affine(A.vk_g1[0]));
        quad_add<Fp>(g_a, *A.a_inputs);
        quad_add<Fp>(g_a, *A.a_aux);
        if (lead) encode_compressed<Fp>(g_a.to_affine(), A.proof);
    } else if (blockIdx.x == 1) {     // B = delta2*s + beta2 + (b2_inputs + b2_aux)   :317-318,336-339
        G2XYZZ g_b = tmp2[0];
        quad_add<Fp2>(g_b, G2XYZZ::from_affine(A.vk_g2[0]));
        quad_add<Fp2>(g_b, *A.b2_inputs);
        quad_add<Fp2>(g_b, *A.b2_aux);
        if (lead) encode_compressed<Fp2>(g_b.to_affine(), A.proof + 48);
    } else {                          // C                                              :319-343
        G1XYZZ g_c = tmp1[1];
        quad_add<Fp>(g_c, tmp1[2]);
        quad_add<Fp>(g_c, tmp1[3]);
        quad_add<Fp>(g_c, tmp1[4]);
        quad_add<Fp>(g_c, tmp1[5]);
        quad_add<Fp>(g_c, *A.h);
        quad_add<Fp>(g_c, *A.l);
        if (lead) encode_compressed<Fp>(g_c.to_affine(), A.proof + 144);
    }
}

// Sharded create_proof: rank partials [world][6 G1 | 2 G2] -> the eight multiexp results.
// One block per multiexp, thread 0 folds the <= 64 partials (8 on one node).
__global__ void prove_fold_partials_kernel(const uint8_t* all, uint32_t world, G1XYZZ* out1, G2XYZZ* out2) {
    if (threadIdx.x >= 4) return;          // one quad per multiexp
    const uint32_t j = blockIdx.x;
    const size_t stride = 6 * sizeof(G1XYZZ) + 2 * sizeof(G2XYZZ);
    if (j < 6) {
        G1XYZZ acc = G1XYZZ::identity();
        for (uint32_t w = 0; w < world; w++) {
            G1XYZZ v = load_struct(reinterpret_cast<const G1XYZZ*>(all + w * stride) + j);
            quad_add<Fp>(acc, v);
        }
        if (threadIdx.x == 0) store_struct(out1 + j, acc);
    } else {
        G2XYZZ acc = G2XYZZ::identity();
        for (uint32_t w = 0; w < world; w++) {
            G2XYZZ v = load_struct(reinterpret_cast<const G2XYZZ*>(all + w * stride + 6 * sizeof(G1XYZZ)) + (j - 6));
            quad_add<Fp2>(acc, v);
        }
        if (threadIdx.x == 0) store_struct(out2 + (j - 6), acc);
    }
}
int prove_fold_partials_launch(bmpc_ctx* ctx, const uint8_t* d_all, uint32_t world, G1XYZZ* out1, G2XYZZ* out2,
                               cudaStream_t st) {
    LAUNCH(ctx, prove_fold_partials_kernel, 8, 32, 0, st, d_all, world, out1, out2);
    return BMPC_OK;
}

int prove_tail_launch(bmpc_ctx* ctx, const ProveTailArgs& args, cudaStream_t st) {
    // scratch for the seven products lives in the context's small device stage (offset 2048)
    G1XYZZ* tmp1 = reinterpret_cast<G1XYZZ*>(ctx->d_stage + 2048);   // 6 x 192 B
    G2XYZZ* tmp2 = reinterpret_cast<G2XYZZ*>(ctx->d_stage + 2048 + 1280);   // 384 B
    LAUNCH(ctx, prove_tail_mul_kernel, 7, 32, 0, st, args, tmp1, tmp2);
    LAUNCH(ctx, prove_tail_kernel, 3, 32, 0, st, args, (const G1XYZZ*)tmp1, (const G2XYZZ*)tmp2);
    return BMPC_OK;
}

}  // namespace bmpc
