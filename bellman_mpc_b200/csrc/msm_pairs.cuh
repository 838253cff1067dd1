// Bucket accumulation as ROUNDS OF INDEPENDENT PAIR ADDITIONS (stage 4 of the MSM pipeline for large
// multiexps; replaces the per-thread tree walk of msm_affine.cuh there).
//
// Same arithmetic as msm_affine.cuh -- affine addition with the inversions batched by Montgomery's
// trick, 6 field products per addition, this is the reference's `buckets[d-1].add_assign_mixed(base)`
// (multiexp.rs:217) re-associated into a pairwise tree -- but the tree is no longer walked by one
// thread per bucket slice.  The (bucket slice, tree level) structure is flattened ON THE DEVICE into
// one flat list of pairs per round:
//
//   round 0   pairs (sorted'[2k], sorted'[2k+1]) of the bucket-sorted entry array itself; every bucket
//             is padded to an even number of entries (pad entry = BMPC_PAIR_PAD: its pair is a copy),
//             operands are gathered from the (window-table) base array
//   round r   pairs listed explicitly (8 B: two indices into the pool of earlier results); the odd
//             element of a slice is CARRIED -- referenced by a later round's list -- never copied and
//             never occupies a lane, so a slice of k points costs exactly k - 1 additions
//   result of pair k of round r -> pool[base_r + k]
//
// so that every thread of every round runs the same straight loop over consecutive list positions:
// no per-thread cursor state, no divergence between lanes, block barriers only around the shared
// inversion, and operand addresses known two iterations ahead.  That is what the previous kernel
// lacked (ncu, profiles/r01d_*: long_scoreboard 3.8 + barrier 1.0 of 14.4 cycles per issue, multiplier
// pipe 66 % busy against 87 % for the XYZZ chain): here the operands of the NEXT pair are staged into
// shared memory with cp.async (LDGSTS, 16 B per request, thread-private slots, conflict-free
// [unit][thread] layout) while the current pair is multiplied.
//
// A block processes chunks of BLK x K consecutive pairs: forward pass (denominators d_k = x2 - x1,
// running product per thread, prefix products to a coalesced global scratch), ONE inversion per
// block per chunk (product tree over the block in shared memory, one thread inverts the root),
// backward pass (1/d_k, slope, x3, y3, coalesced store of the sum).  P + P (tangent), P + (-P),
// identity operands and pad entries are classified per pair exactly as in msm_affine.cuh.
//
// The pair lists are built from the task descriptors by pair_build_kernel (msm_sort_kernels.cuh);
// `pair_build_task` below is the per-slice rule, shared with the host test (tests/host_check).
#pragma once
#include <stddef.h>

#include "msm_affine.cuh"
#include "msm_pair_lists.h"

namespace bmpc {

// ------------------------------------------------------------------ the addition itself
// forward step: denominator of P + Q from the x coordinates alone.  Returns 0: multiply d into the
// running product; 1: nothing to multiply (pad entry); 2: rare -- equal x or a zero x (identity,
// tangent or opposite points): the caller fetches both points and asks aff_classify
template <class F>
BMPC_HD int pair_fwd_quick(const F& x1, const F& x2, bool single, F& d) {
    if (single) return 1;
    d = x2 - x1;
    return (d.is_zero() || x1.is_zero() || x2.is_zero()) ? 2 : 0;
}
// backward step: `inv` = 1 / (d_0 .. d_k) on entry, 1 / (d_0 .. d_{k-1}) on exit; pk = d_0 .. d_{k-1}
template <class F>
BMPC_HD Affine<F> pair_bwd_add(const Affine<F>& P, const Affine<F>& Q, const F& pk, F& inv) {
    F d;
    const int kind = aff_classify<F>(P, Q, d);
    if (kind == 2) return aff_trivial<F>(P, Q);
    F dinv = inv * pk;
    inv = inv * d;
    F num = Q.y - P.y;
    if (kind == 1) {                       // rare: tangent
        F px = P.x;
        F xx = F::mul_cold(px, px);
        num = xx.dbl() + xx;
    }
    F lam = num * dinv;
    Affine<F> R;
    R.x = lam.sqr() - P.x - Q.x;
    R.y = lam * (P.x - R.x) - P.y;
    return R;
}

#if defined(__CUDACC__)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <class F>
struct PairArgs {
    const Affine<F>* tables;     // round 0 operands (gathered through the sorted entries)
    Affine<F>* out;              // pool of round results
    const uint2* list;           // round r >= 1: (a, b) pool indices; round 0: sorted' viewed as pairs
    const uint32_t* count_p;     // device: number of pairs (round 0: the padded entry count, shifted)
    uint32_t count_shift;
    uint32_t out_base;           // pool index of this round's first result
    F* pre;                      // gridDim.x * BLK * kmax prefix products (block-private regions)
    uint32_t kmax;               // most pairs per thread per chunk
};

// Shared-memory staging: `STAGES` slots per thread, each 2 points (+ the prefix product when PK),
// laid out [stage][16-byte unit][thread] so that a warp's accesses are 512 contiguous bytes.
template <class F, int BLK, bool PK>
struct PairStage {
    static constexpr int PU = sizeof(Affine<F>) / 16;       // units per point
    static constexpr int XU = sizeof(F) / 16;               // units per coordinate
    static constexpr int UNITS = 2 * PU + (PK ? XU : 0);
    static constexpr size_t BYTES = (size_t)2 * UNITS * BLK * 16;
    uint4* base;
    __device__ __forceinline__ uint4* at(int stage, int unit) const {
        return base + ((size_t)(stage * UNITS + unit) * BLK + threadIdx.x);
    }
    template <int N>
    __device__ __forceinline__ void issue(int stage, int unit0, const void* g) const {
        const uint4* s = reinterpret_cast<const uint4*>(g);
#pragma unroll
        for (int u = 0; u < N; u++) cp_async16(at(stage, unit0 + u), s + u);
    }
    template <class T>
    __device__ __forceinline__ T read(int stage, int unit0) const {
        T r;
        uint4* w = reinterpret_cast<uint4*>(&r);
#pragma unroll
        for (int u = 0; u < (int)(sizeof(T) / 16); u++) w[u] = *at(stage, unit0 + u);
        return r;
    }
};

// One round.  Persistent grid (blocks resident at once: MINB per SM); dynamic shared memory =
// max(PairStage::BYTES, 4 * BLK * sizeof(F)) (the inversion tree reuses the staging area).
template <class F, bool R0, int BLK, int MINB, bool PK>
__global__ void __launch_bounds__(BLK, MINB)
msm_pair_round_kernel(PairArgs<F> a) {
    extern __shared__ uint4 pair_smem[];
    typedef PairStage<F, BLK, PK> Stage;
    Stage stg;
    stg.base = pair_smem;
    BlockCoop<F> coop;
    coop.B = BLK;
    coop.P = reinterpret_cast<F*>(pair_smem);
    coop.I = coop.P + 2 * BLK;
    const uint32_t tid = threadIdx.x;
    const uint32_t P = *a.count_p >> a.count_shift;
    if (P == 0) return;
    // chunks: every block runs `nch` chunks of BLK x K consecutive pairs, K <= kmax chosen so that the
    // last wave is as full as the first
    const uint64_t wave = (uint64_t)gridDim.x * BLK;
    const uint32_t nch = (uint32_t)((P + wave * a.kmax - 1) / (wave * a.kmax));
    const uint32_t K = (uint32_t)((P + wave * nch - 1) / (wave * nch));
    F* pre = a.pre + (size_t)blockIdx.x * BLK * a.kmax + tid;      // element j at pre[j * BLK]
    const Affine<F>* src = R0 ? a.tables : a.out;
    for (uint32_t c = 0; c < nch; c++) {
        const uint64_t cbase = ((uint64_t)c * gridDim.x + blockIdx.x) * (uint64_t)K * BLK;
        if (cbase >= P) break;                                         // block-uniform
        uint32_t kt = 0;
        if (cbase + tid < P) {
            uint64_t left = (P - cbase - tid + BLK - 1) / BLK;
            kt = left < K ? (uint32_t)left : K;
        }
        const uint2* lst = a.list + cbase + tid;                        // pair j at lst[j * BLK]
        // ---------------------------------------------------------------- forward
        F acc = F::one();
        {
            uint2 e0 = make_uint2(0, 0), e1 = make_uint2(0, 0);
            if (kt > 0) e0 = __ldg(lst);
            if (kt > 1) e1 = __ldg(lst + BLK);
            if (kt > 0) {
                const uint32_t ia = R0 ? (e0.x & 0x7fffffffu) : e0.x;
                const uint32_t ib = R0 ? (e0.y == BMPC_PAIR_PAD ? ia : (e0.y & 0x7fffffffu)) : e0.y;
                stg.template issue<Stage::XU>(0, 0, &src[ia].x);
                stg.template issue<Stage::XU>(0, Stage::XU, &src[ib].x);
            }
            cp_async_commit();
#pragma unroll 1
            for (uint32_t j = 0; j < kt; j++) {
                uint2 e2 = make_uint2(0, 0);
                if (j + 2 < kt) e2 = __ldg(lst + (size_t)(j + 2) * BLK);
                if (j + 1 < kt) {
                    const uint32_t ia = R0 ? (e1.x & 0x7fffffffu) : e1.x;
                    const uint32_t ib = R0 ? (e1.y == BMPC_PAIR_PAD ? ia : (e1.y & 0x7fffffffu)) : e1.y;
                    stg.template issue<Stage::XU>((j + 1) & 1, 0, &src[ia].x);
                    stg.template issue<Stage::XU>((j + 1) & 1, Stage::XU, &src[ib].x);
                }
                cp_async_commit();
                cp_async_wait<1>();
                const int s = j & 1;
                F x1 = stg.template read<F>(s, 0);
                F x2 = stg.template read<F>(s, Stage::XU);
                const bool single = R0 && e0.y == BMPC_PAIR_PAD;
                F d;
                const int q = pair_fwd_quick<F>(x1, x2, single, d);
                bool use = q == 0;
                if (q == 2) {                                          // rare
                    Affine<F> Pp = aff_ld(src + (R0 ? (e0.x & 0x7fffffffu) : e0.x));
                    Affine<F> Qq = aff_ld(src + (R0 ? (e0.y & 0x7fffffffu) : e0.y));
                    if (R0 && (e0.x & 0x80000000u)) Pp.y = Pp.y.neg();
                    if (R0 && (e0.y & 0x80000000u)) Qq.y = Qq.y.neg();
                    use = aff_classify<F>(Pp, Qq, d) != 2;
                }
                aff_st(pre + (size_t)j * BLK, acc);
                if (use) acc = acc * d;
                e0 = e1;
                e1 = e2;
            }
            cp_async_wait<0>();
        }
        // ---------------------------------------------------------------- one inversion per block
        __syncthreads();                   // the tree reuses the staging area
        F inv = coop.invert(acc);
        __syncthreads();
        // ---------------------------------------------------------------- backward
        {
            uint2 e0 = make_uint2(0, 0), e1 = make_uint2(0, 0);
            if (kt > 0) e0 = __ldg(lst + (size_t)(kt - 1) * BLK);
            if (kt > 1) e1 = __ldg(lst + (size_t)(kt - 2) * BLK);
            auto issue_bwd = [&](int s, const uint2& e, uint32_t j) {
                const uint32_t ia = R0 ? (e.x & 0x7fffffffu) : e.x;
                const uint32_t ib = R0 ? (e.y == BMPC_PAIR_PAD ? ia : (e.y & 0x7fffffffu)) : e.y;
                stg.template issue<Stage::PU>(s, 0, src + ia);
                stg.template issue<Stage::PU>(s, Stage::PU, src + ib);
                if (PK) stg.template issue<Stage::XU>(s, 2 * Stage::PU, pre + (size_t)j * BLK);
            };
            if (kt > 0) issue_bwd((kt - 1) & 1, e0, kt - 1);
            cp_async_commit();
#pragma unroll 1
            for (uint32_t j = kt; j-- > 0;) {
                uint2 e2 = make_uint2(0, 0);
                if (j >= 2) e2 = __ldg(lst + (size_t)(j - 2) * BLK);
                if (j >= 1) issue_bwd((j - 1) & 1, e1, j - 1);
                cp_async_commit();
                F pk;
                if (!PK) pk = aff_ld(pre + (size_t)j * BLK);
                cp_async_wait<1>();
                const int s = j & 1;
                Affine<F> Pp = stg.template read<Affine<F>>(s, 0);
                Affine<F> Qq = stg.template read<Affine<F>>(s, Stage::PU);
                if (PK) pk = stg.template read<F>(s, 2 * Stage::PU);
                if (R0) {
                    if (e0.x & 0x80000000u) Pp.y = Pp.y.neg();
                    if (e0.y == BMPC_PAIR_PAD) Qq = Affine<F>::identity();
                    else if (e0.y & 0x80000000u) Qq.y = Qq.y.neg();
                }
                Affine<F> R = pair_bwd_add<F>(Pp, Qq, pk, inv);
                aff_st(a.out + (size_t)a.out_base + cbase + (size_t)j * BLK + tid, R);
                e0 = e1;
                e1 = e2;
            }
            cp_async_wait<0>();
        }
        __syncthreads();                   // staging area free again before the next chunk's tree / loads
    }
}

// slice sums -> the XYZZ partial-sum slots the bucket reduction reads
template <class F>
__global__ void __launch_bounds__(128)
msm_pair_collect_kernel(const Affine<F>* out, const uint32_t* fin, const uint4* desc, const uint32_t* ntasks_p,
                        XYZZ<F>* partials) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= *ntasks_p) return;
    const uint32_t f = fin[t];
    XYZZ<F> v = XYZZ<F>::identity();
    if (f != BMPC_PAIR_NONE) v = XYZZ<F>::from_affine(aff_ld(out + f));
    aff_st(partials + __ldg(desc + t).z, v);
}
#endif

}  // namespace bmpc
