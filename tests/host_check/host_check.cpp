// Host-side build of the device arithmetic headers (carry flag emulated) so the
// algorithms can be checked bit-for-bit against the big-integer oracle without a GPU.
// Test infrastructure only.
#include "../../bellman_mpc_b200/csrc/curve.cuh"
#include <string.h>
using namespace bmpc;

extern "C" {
// op: 0 mul, 1 add, 2 sub, 3 neg, 4 inv, 5 to_mont, 6 from_mont, 7 sqr, 8 inv_fermat
void hc_fr_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* r) {
    Fr x, y, z; memcpy(x.l, a, 32); memcpy(y.l, b, 32);
    switch (op) { case 0: z = x * y; break; case 1: z = x + y; break; case 2: z = x - y; break;
        case 3: z = x.neg(); break; case 4: z = x.inv(); break; case 5: z = x.to_mont(); break;
        case 6: z = x.from_mont(); break; case 8: z = x.inv_fermat(); break; default: z = x.sqr_redc(); }
    memcpy(r, z.l, 32);
}
void hc_fp_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* r) {
    Fp x, y, z; memcpy(x.l, a, 48); memcpy(y.l, b, 48);
    switch (op) { case 0: z = x * y; break; case 1: z = x + y; break; case 2: z = x - y; break;
        case 3: z = x.neg(); break; case 4: z = x.inv(); break; case 5: z = x.to_mont(); break;
        case 6: z = x.from_mont(); break; case 8: z = x.inv_fermat(); break; default: z = x.sqr_redc(); }
    memcpy(r, z.l, 48);
}
void hc_fp2_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* r) {
    Fp2 x, y, z; memcpy(&x, a, 96); memcpy(&y, b, 96);
    switch (op) { case 0: z = x * y; break; case 1: z = x + y; break; case 2: z = x - y; break;
        case 3: z = x.neg(); break; case 4: z = x.inv(); break; default: z = x.sqr(); }
    memcpy(r, &z, 96);
}
// G1 ops on affine Montgomery points (24 words each, (0,0) = identity)
// op: 0 add (via XYZZ add_affine), 1 full add (XYZZ+XYZZ with non-trivial Z), 2 double, 3 scalar mul by k[8]
void hc_g1_op(int op, const uint32_t* a, const uint32_t* b, const uint32_t* k, uint32_t* r) {
    G1Affine p, q; memcpy(&p, a, 96); memcpy(&q, b, 96);
    G1XYZZ acc = G1XYZZ::from_affine(p);
    if (op == 0) acc.add_affine(q);
    else if (op == 1) {
        // give both operands non-trivial denominators: P = (2P - P'), Q likewise
        G1XYZZ P2 = acc.dbl(); P2.add_affine(p.neg());
        G1XYZZ Q = G1XYZZ::from_affine(q); G1XYZZ Q2 = Q.dbl(); Q2.add_affine(q.neg());
        P2.add(Q2); acc = P2;
    } else if (op == 2) acc = acc.dbl();
    else acc = acc.mul(k, 8);
    G1Affine o = acc.to_affine(); memcpy(r, &o, 96);
}
void hc_g2_op(int op, const uint32_t* a, const uint32_t* b, const uint32_t* k, uint32_t* r) {
    G2Affine p, q; memcpy(&p, a, 192); memcpy(&q, b, 192);
    G2XYZZ acc = G2XYZZ::from_affine(p);
    if (op == 0) acc.add_affine(q);
    else if (op == 1) {
        G2XYZZ P2 = acc.dbl(); P2.add_affine(p.neg());
        G2XYZZ Q = G2XYZZ::from_affine(q); G2XYZZ Q2 = Q.dbl(); Q2.add_affine(q.neg());
        P2.add(Q2); acc = P2;
    } else if (op == 2) acc = acc.dbl();
    else acc = acc.mul(k, 8);
    G2Affine o = acc.to_affine(); memcpy(r, &o, 192);
}
}
