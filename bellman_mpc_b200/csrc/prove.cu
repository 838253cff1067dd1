// Tail of create_proof (reference: src/groth16/prover.rs:309-349): assemble A, B, C from the
// eight multiexp results and the verifying-key points, normalise, and write the 192-byte
// compressed proof (Proof::write, src/groth16/mod.rs:42-48).  One thread: ~10 scalar
// multiplications and three inversions -- negligible next to the MSMs, kept on the device so
// that no group arithmetic runs on the host.
#include "encode.cuh"
#include "internal.h"

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

__global__ void prove_tail_kernel(ProveTailArgs A) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Fr r = A.rs[0], s = A.rs[1];
    Fr rs = r * s;                                   // prover.rs:322-323
    Fr rc = r.from_mont(), sc = s.from_mont(), rsc = rs.from_mont();
    G1XYZZ alpha = G1XYZZ::from_affine(A.vk_g1[0]);
    G1XYZZ beta1 = G1XYZZ::from_affine(A.vk_g1[1]);
    G1XYZZ delta1 = G1XYZZ::from_affine(A.vk_g1[2]);
    G2XYZZ beta2 = G2XYZZ::from_affine(A.vk_g2[0]);
    G2XYZZ delta2 = G2XYZZ::from_affine(A.vk_g2[1]);

    G1XYZZ g_a = delta1.mul(rc.l, 8);                // :315-316
    g_a.add(alpha);
    G2XYZZ g_b = delta2.mul(sc.l, 8);                // :317-318
    g_b.add(beta2);
    G1XYZZ g_c = delta1.mul(rsc.l, 8);               // :319-327
    g_c.add(alpha.mul(sc.l, 8));
    g_c.add(beta1.mul(rc.l, 8));

    G1XYZZ a_answer = *A.a_inputs;                   // :328-332
    a_answer.add(*A.a_aux);
    g_a.add(a_answer);
    g_c.add(a_answer.mul(sc.l, 8));

    G1XYZZ b1_answer = *A.b1_inputs;                 // :334-337
    b1_answer.add(*A.b1_aux);
    G2XYZZ b2_answer = *A.b2_inputs;
    b2_answer.add(*A.b2_aux);

    g_b.add(b2_answer);                              // :339-343
    g_c.add(b1_answer.mul(rc.l, 8));
    g_c.add(*A.h);
    g_c.add(*A.l);

    encode_compressed<Fp>(g_a.to_affine(), A.proof);         // :345-349 + Proof::write
    encode_compressed<Fp2>(g_b.to_affine(), A.proof + 48);
    encode_compressed<Fp>(g_c.to_affine(), A.proof + 144);
}

int prove_tail_launch(bmpc_ctx* ctx, const ProveTailArgs& args, cudaStream_t st) {
    LAUNCH(ctx, prove_tail_kernel, 1, 1, 0, st, args);
    return BMPC_OK;
}

}  // namespace bmpc
