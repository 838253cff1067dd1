// Batched-affine bucket accumulation (stage 4 of the MSM pipeline, see group_kernels.cuh).
//
// The XYZZ accumulate kernel spends 10 field products per `bucket += base` (madd-2008-s).  In
// affine coordinates the same addition is  lambda = (y2 - y1) / (x2 - x1),  x3 = lambda^2 - x1 - x2,
// y3 = lambda (x1 - x3) - y1: 3 products and one inversion; with Montgomery's trick K
// independent inversions cost 3 K products + ONE inversion, i.e. 6 products per addition.
// Independent additions come from summing every bucket slice as a pairwise TREE: round 0 adds
// points (2i, 2i+1) of the slice (read from the bases through `sorted`), round r adds the
// results of round r-1, until one point is left.  All pairs of a round are independent.
//
// One thread owns a JOB of BMPC_AFF_G consecutive tasks (slices of <= L points; tasks are ordered
// by size, so the slices of a job -- and the jobs of a warp -- have nearly equal lengths) and walks
// the pairs of all its slices in chunks of BMPC_AFF_K: forward pass (denominators, running
// product kept in thread-private local memory), one inversion, backward pass (recover each
// 1/d, finish the addition, store the sum).  Intermediate points live in a thread-private
// scratch region of HBM: buffer A (ceil(L/2) points per slice) and buffer B (half of that),
// used alternately as destination.
//
// The inversion itself is shared by the whole thread block (BlockCoop): the 128 chunk products
// go up a product tree in shared memory, ONE thread inverts the root (binary extended Euclid,
// ~50 K data-dependent instructions: run per thread it diverges across the warp and cost as
// much as the additions it was meant to save -- measured 76.8 ms vs 74.8 ms for XYZZ at 2^24), and
// the inverses come back down the tree: 14 products per thread per chunk instead of an
// inversion.  All loop bounds are made block-uniform through Coop::any so every thread reaches
// every barrier.
//
// Complete: identity operands, P + P (doubling: lambda = 3 x^2 / 2 y with its own denominator
// in the same batch) and P + (-P) are classified per pair and never put a zero into the product.
// This is the reference's `buckets[d-1].add_assign_mixed(base)` (multiexp.rs:217) re-associated;
// group addition is associative and the final affine result canonical, so the bytes are the same.
//
// The per-job body is BMPC_HD so that tests/host_check can run it on the CPU against the XYZZ
// chain.
#pragma once
#include <stddef.h>

#include "curve.cuh"

namespace bmpc {

#ifndef BMPC_AFF_G
#define BMPC_AFF_G 8      // tasks (bucket slices) per job
#endif
#ifndef BMPC_AFF_K
#define BMPC_AFF_K 128    // additions per inversion
#endif

template <class T>
BMPC_HD T aff_ld(const T* p) {          // coherent load (scratch written by this thread)
#if defined(__CUDA_ARCH__)
    T r;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4* w = reinterpret_cast<uint4*>(&r);
#pragma unroll
    for (int j = 0; j < (int)(sizeof(T) / 16); j++) w[j] = q[j];
    return r;
#else
    return *p;
#endif
}
template <class T>
BMPC_HD T aff_ldg(const T* p) {         // read-only path (bases, sorted)
#if defined(__CUDA_ARCH__)
    T r;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4* w = reinterpret_cast<uint4*>(&r);
#pragma unroll
    for (int j = 0; j < (int)(sizeof(T) / 16); j++) w[j] = __ldg(q + j);
    return r;
#else
    return *p;
#endif
}
template <class T>
BMPC_HD void aff_st(T* p, const T& v) {
#if defined(__CUDA_ARCH__)
    uint4* q = reinterpret_cast<uint4*>(p);
    const uint4* w = reinterpret_cast<const uint4*>(&v);
#pragma unroll
    for (int j = 0; j < (int)(sizeof(T) / 16); j++) q[j] = w[j];
#else
    *p = v;
#endif
}
BMPC_HD uint32_t aff_ldg_u32(const uint32_t* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

template <class F>
struct AffJob {
    const Affine<F>* bases;     // CRS points (all tables)
    const uint32_t* sorted;     // base index | sign << 31, grouped by bucket
    Affine<F>* bufA;            // G x HA points, thread-private
    Affine<F>* bufB;            // G x HB points
    uint32_t HA, HB;
    uint32_t G;                 // slices in this job's geometry, <= BMPC_AFF_G
    uint32_t start[BMPC_AFF_G]; // first `sorted` entry of each slice
    uint32_t len[BMPC_AFF_G];   // points left in each slice (0: no task)
    uint32_t slot[BMPC_AFF_G];  // partial-sum slot of each slice
};

// point `idx` of slice g as the current round sees it
template <class F>
BMPC_HD Affine<F> aff_fetch(const AffJob<F>& J, bool r0, const Affine<F>* src, uint32_t sH, uint32_t g,
                            uint32_t idx) {
    if (r0) {
        uint32_t e = aff_ldg_u32(J.sorted + J.start[g] + idx);
        Affine<F> p = aff_ldg(J.bases + (e & 0x7fffffffu));
        if (e & 0x80000000u) p.y = p.y.neg();     // -(0,0) = (0,0): the identity stays the identity
        return p;
    }
    return aff_ld(src + (size_t)g * sH + idx);
}
// its x coordinate only (forward pass)
template <class F>
BMPC_HD F aff_fetch_x(const AffJob<F>& J, bool r0, const Affine<F>* src, uint32_t sH, uint32_t g, uint32_t idx) {
    if (r0) {
        uint32_t e = aff_ldg_u32(J.sorted + J.start[g] + idx);
        return aff_ldg(&J.bases[e & 0x7fffffffu].x);
    }
    return aff_ld(&src[(size_t)g * sH + idx].x);
}

// Round 0 reads its operands through `sorted`: a dependent pair of loads (index, then point).
// Both passes load the index entries of the NEXT pair before they start multiplying on the
// current one, so only the point loads are exposed (the index hop was 6.7 % of the stall samples).
struct AffEntries { uint32_t x, y; };
template <class F>
BMPC_HD AffEntries aff_pair_entries(const AffJob<F>& J, uint32_t g, uint32_t i) {
    AffEntries e;
    e.x = aff_ldg_u32(J.sorted + J.start[g] + 2 * i);
    e.y = aff_ldg_u32(J.sorted + J.start[g] + 2 * i + 1);
    return e;
}
// the same with the slice's first entry already in a register
BMPC_HD AffEntries aff_pair_entries_at(const uint32_t* sorted, uint32_t start, uint32_t i) {
    AffEntries e;
    e.x = aff_ldg_u32(sorted + start + 2 * i);
    e.y = aff_ldg_u32(sorted + start + 2 * i + 1);
    return e;
}
template <class F>
BMPC_HD Affine<F> aff_fetch_e(const AffJob<F>& J, bool r0, uint32_t e, const Affine<F>* src, uint32_t sH, uint32_t g,
                              uint32_t idx) {
    if (r0) {
        Affine<F> p = aff_ldg(J.bases + (e & 0x7fffffffu));
        if (e & 0x80000000u) p.y = p.y.neg();
        return p;
    }
    return aff_ld(src + (size_t)g * sH + idx);
}
template <class F>
BMPC_HD F aff_fetch_x_e(const AffJob<F>& J, bool r0, uint32_t e, const Affine<F>* src, uint32_t sH, uint32_t g,
                        uint32_t idx) {
    if (r0) return aff_ldg(&J.bases[e & 0x7fffffffu].x);
    return aff_ld(&src[(size_t)g * sH + idx].x);
}

// 0: chord (d = x2 - x1), 1: tangent (d = 2 y1), 2: no inversion needed
template <class F>
BMPC_HD int aff_classify(const Affine<F>& P, const Affine<F>& Q, F& d) {
    if (P.is_identity() || Q.is_identity()) return 2;
    d = Q.x - P.x;
    if (!d.is_zero()) return 0;
    if (P.y == Q.y) {
        d = P.y.dbl();
        return d.is_zero() ? 2 : 1;      // y == 0: 2-torsion, 2 P = identity
    }
    return 2;                             // P == -Q
}
template <class F>
BMPC_HD Affine<F> aff_trivial(const Affine<F>& P, const Affine<F>& Q) {
    if (P.is_identity()) return Q;
    if (Q.is_identity()) return P;
    return Affine<F>::identity();
}

// How the threads that run jobs side by side cooperate.  Solo: nobody else (host tests).
struct SoloCoop {
    BMPC_HD bool any(bool p) const { return p; }
    template <class F>
    BMPC_HD F invert(const F& acc) const { return acc.inv(); }
};

// Sums every slice of the job; leaves slice g's total (affine) in partials[slot[g]] as XYZZ.
// `pre`: thread-private array of BMPC_AFF_K field elements.
template <class F, class Coop>
BMPC_HD void aff_run_job(AffJob<F>& J, F* pre, uint32_t K, XYZZ<F>* partials, const Coop& coop) {
    const uint32_t G = J.G;
    bool r0 = true;
    const Affine<F>* src = nullptr;
    uint32_t sH = 0;
    Affine<F>* dst = J.bufA;
    uint32_t dH = J.HA;
    for (;;) {
        uint32_t maxlen = 0;
#pragma unroll 1
        for (uint32_t g = 0; g < G; g++) maxlen = J.len[g] > maxlen ? J.len[g] : maxlen;
        if (!coop.any(r0 || maxlen > 1)) break;
        // cursor over the pairs (g, i), i < len[g] / 2.  The job descriptor lives in local memory (it is
        // indexed by g), so the current slice's pair count and first entry are kept in registers and
        // re-read only when the cursor moves to another slice: the per-pair `J.len[g]` / `J.start[g]`
        // loads were 40 % of the kernel's long-scoreboard stall samples (ncu source view, round 2).
        uint32_t g = 0, i = 0;
        while (g < G && (J.len[g] >> 1) == 0) g++;
        uint32_t half = 0, st0 = 0;
        if (g < G) { half = J.len[g] >> 1; st0 = J.start[g]; }
        while (coop.any(g < G)) {
            // ---- forward: denominators and their running product
            F acc = F::one();
            uint32_t cnt = 0;
#pragma unroll 1
            AffEntries e = {0u, 0u};
            if (r0 && g < G) e = aff_pair_entries_at(J.sorted, st0, i);
#pragma unroll 1
            while (cnt < K && g < G) {
                const uint32_t gc = g, ic = i;          // the current pair; (g, i) moves on to the next
                F x1 = aff_fetch_x_e<F>(J, r0, e.x, src, sH, gc, 2 * ic);
                F x2 = aff_fetch_x_e<F>(J, r0, e.y, src, sH, gc, 2 * ic + 1);
                const AffEntries ec = e;
                i++;
                if (i >= half) {
                    do { g++; } while (g < G && (J.len[g] >> 1) == 0);
                    i = 0;
                    if (g < G) { half = J.len[g] >> 1; st0 = J.start[g]; }
                }
                if (r0 && g < G) e = aff_pair_entries_at(J.sorted, st0, i);
                F d = x2 - x1;
                bool use = true;
                if (d.is_zero() || x1.is_zero() || x2.is_zero()) {       // rare
                    Affine<F> P = aff_fetch_e<F>(J, r0, ec.x, src, sH, gc, 2 * ic);
                    Affine<F> Q = aff_fetch_e<F>(J, r0, ec.y, src, sH, gc, 2 * ic + 1);
                    use = aff_classify<F>(P, Q, d) != 2;
                }
                aff_st(pre + cnt, acc);
                if (use) acc = acc * d;
                cnt++;
            }
            F inv = coop.invert(acc);
            // ---- backward: walk the same pairs in reverse
            uint32_t g2 = g, i2 = i, st2 = st0;
            if (cnt) {                                   // step back onto the chunk's last pair
                if (i2 == 0) {
                    do { g2--; } while ((J.len[g2] >> 1) == 0);
                    i2 = J.len[g2] >> 1;
                    st2 = J.start[g2];
                }
                i2--;
                if (r0) e = aff_pair_entries_at(J.sorted, st2, i2);
            }
#pragma unroll 1
            for (uint32_t k = cnt; k-- > 0;) {
                const uint32_t gc = g2, ic = i2;
                Affine<F> P = aff_fetch_e<F>(J, r0, e.x, src, sH, gc, 2 * ic);
                Affine<F> Q = aff_fetch_e<F>(J, r0, e.y, src, sH, gc, 2 * ic + 1);
                if (k > 0) {                             // the pair before this one, and its entries
                    if (i2 == 0) {
                        do { g2--; } while ((J.len[g2] >> 1) == 0);
                        i2 = J.len[g2] >> 1;
                        st2 = J.start[g2];
                    }
                    i2--;
                    if (r0) e = aff_pair_entries_at(J.sorted, st2, i2);
                }
                F d;
                int kind = aff_classify<F>(P, Q, d);
                Affine<F> R;
                if (kind == 2) {
                    R = aff_trivial<F>(P, Q);
                } else {
                    F pk = aff_ld(pre + k);
                    F dinv = inv * pk;              // 1 / d_k
                    inv = inv * d;                  // 1 / (d_0 ... d_{k-1})
                    F num = Q.y - P.y;
                    if (kind == 1) {               // rare: keep the call's operands branch-local
                        F px = P.x;
                        F xx = F::mul_cold(px, px);
                        num = xx.dbl() + xx;
                    }
                    F lam = num * dinv;
                    R.x = lam.sqr() - P.x - Q.x;
                    R.y = lam * (P.x - R.x) - P.y;
                }
                aff_st(dst + (size_t)gc * dH + ic, R);
            }
        }
        // odd point of each slice moves up unchanged; lengths halve
#pragma unroll 1
        for (uint32_t s = 0; s < G; s++) {
            uint32_t l = J.len[s];
            if (l & 1u) {
                Affine<F> P = aff_fetch<F>(J, r0, src, sH, s, l - 1);
                aff_st(dst + (size_t)s * dH + (l >> 1), P);
            }
            J.len[s] = (l + 1) >> 1;
        }
        // next round reads what this one wrote
        src = dst;
        sH = dH;
        if (dst == J.bufA) { dst = J.bufB; dH = J.HB; } else { dst = J.bufA; dH = J.HA; }
        r0 = false;
    }
#pragma unroll 1
    for (uint32_t s = 0; s < G; s++) {
        if (!J.len[s]) continue;
        Affine<F> P = aff_ld(src + (size_t)s * sH);
        XYZZ<F> v = XYZZ<F>::from_affine(P);
        aff_st(partials + J.slot[s], v);
    }
}

#if defined(__CUDACC__)
// Block-wide batched inversion.  P and I: 2 * B field elements of shared memory each (binary
// trees, node n has children 2n and 2n+1, leaves at B + thread).
template <class F>
struct BlockCoop {
    // The 14 products per thread per chunk of the product tree.  G2: out of line -- fully inlined the
    // G2 kernel is 265 KB of code and `no_instruction` is its top stall (ncu, profiles/r02_*_g2_*: 4.4
    // cycles per issue); without the tree's two Fp2 sites it is 197 KB and 4.5 % faster (2^21 points:
    // 31.5 -> 30.1 ms; the forward pass' product out of line as well: 32.4 ms, worse).  G1 (108 KB, no
    // instruction starvation) keeps them inlined: 0.3 % faster.
    // (operands copied to registers first: multiplied straight out of shared memory the limbs are loaded
    // in the middle of the carry chains and the lo/hi halves no longer pair into IMAD.WIDE -- see mul_cold)
    static __device__ __forceinline__ F mul_rare(const F& a, const F& b) {
        if constexpr (sizeof(F) > sizeof(Fp)) return F::mul_cold(a, b);
        else {
            const F x = aff_ld(&a), y = aff_ld(&b);
            return x * y;
        }
    }
    F* P;
    F* I;
    uint32_t B;   // threads per block, power of two
    __device__ __forceinline__ bool any(bool p) const { return __syncthreads_or(p) != 0; }
    __device__ __forceinline__ F invert(const F& acc) const {
        const uint32_t j = threadIdx.x;
        P[B + j] = acc;
        __syncthreads();
#pragma unroll 1
        for (uint32_t w = B >> 1; w >= 1; w >>= 1) {
            if (j < w) P[w + j] = mul_rare(P[2 * (w + j)], P[2 * (w + j) + 1]);
            __syncthreads();
        }
        if (j == 0) I[1] = P[1].inv();
        __syncthreads();
#pragma unroll 1
        for (uint32_t w = 1; w < B; w <<= 1) {
            if (j < 2 * w) {
                uint32_t c = 2 * w + j;
                I[c] = mul_rare(I[c >> 1], P[c ^ 1u]);
            }
            __syncthreads();
        }
        return I[B + j];
    }
};

#define BMPC_AFF_BLOCK 128   // largest block; the launch picks blockDim (power of two)
// Persistent grid: thread t runs jobs t, t + T, t + 2T, ... (T = all threads of the grid);
// scratch holds T x G x (HA + HB) points.  G (slices per job) is chosen by the host so that the
// grid is full: big MSMs run G = BMPC_AFF_G, small ones trade amortisation for parallelism.
// Dynamic shared memory: 4 * blockDim.x * sizeof(F).  K = additions per inversion per thread.
template <class F, int K, int MINB, int BLK = BMPC_AFF_BLOCK>
__global__ void __launch_bounds__(BLK, MINB)
msm_accumulate_affine_kernel(const Affine<F>* bases, const uint32_t* sorted, const uint4* desc,
                             const uint32_t* ntasks_p, XYZZ<F>* partials, Affine<F>* scratch, uint32_t HA,
                             uint32_t HB, uint32_t G, uint32_t whole_waves) {
    extern __shared__ uint4 aff_smem[];
    BlockCoop<F> coop;
    coop.B = blockDim.x;
    coop.P = reinterpret_cast<F*>(aff_smem);
    coop.I = coop.P + 2 * blockDim.x;
    const uint32_t ntasks = *ntasks_p;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, T = gridDim.x * blockDim.x;
    // Jobs are handed out in waves of T threads.  ceil(ntasks / G) jobs of G slices leave the last
    // wave partly empty (2^19 slices over 75776 threads with G = 4: 1.73 waves, the kernel runs as
    // long as 2 full ones).  With `whole_waves` the SAME slices are dealt over a whole number of
    // waves instead: njobs = waves * T, so a job carries G or fewer slices (the size classes it
    // misses are the smallest ones) and every thread gets ntasks / T slices' worth of points.
    // Measured (G1): 2^21 points 9.67 -> 9.04 ms, 2^22 18.17 -> 16.93 ms, 2^24 unchanged (3.95 waves).
    uint32_t njobs = (ntasks + G - 1) / G;
    if (whole_waves) njobs = ((njobs + T - 1) / T) * T;
    __align__(16) F pre[K];
    AffJob<F> J;
    J.bases = bases;
    J.sorted = sorted;
    J.bufA = scratch + (size_t)tid * G * (HA + HB);
    J.bufB = J.bufA + (size_t)G * HA;
    J.HA = HA;
    J.HB = HB;
    J.G = G;
    // block-uniform trip count: a thread without a job still walks the barriers with empty slices
    for (uint32_t job = tid; job - threadIdx.x < njobs; job += T) {
#pragma unroll 1
        for (uint32_t g = 0; g < G; g++) {
            // Tasks are ordered by size; job j takes tasks j, j + njobs, j + 2 njobs, ...: one slice
            // from each size class, so all jobs carry (nearly) the same number of points and the
            // lanes of a warp see similar slice lengths.  (Consecutive tasks per job made the first
            // jobs twice as long as the average: a half-empty GPU for the second half of the kernel.)
            uint32_t t = g * njobs + job;
            uint4 d = make_uint4(0, 0, 0, 0);
            if (job < njobs && t < ntasks) d = __ldg(desc + t);   // {first entry, length, partial slot, bucket}
            J.start[g] = d.x;
            J.len[g] = d.y;
            J.slot[g] = d.z;
        }
        aff_run_job<F>(J, pre, (uint32_t)K, partials, coop);
    }
}
#endif

}  // namespace bmpc
