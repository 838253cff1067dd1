// Pippenger bucket multi-scalar multiplication for BLS12-381 G1 / G2.
//
// Replaces `multiexp` / `multiexp_inner` (reference: src/multiexp.rs:159-281).  The reference
// runs one CPU task per c-bit window, each scanning all n (exponent, density) pairs and adding
// bases into 2^c - 1 projective buckets (:191-223), then sums buckets by parts (:229-233) and
// folds windows top-down with c doublings (:244-249).  Here:
//
//   1. msm_count_kernel   one thread per exponent position: density bit -> base index by
//                         prefix popcount (SURVEY 8a'/1), signed c-bit digits (halves the
//                         bucket count; negation of an affine point is free), histogram of
//                         (window, |digit|) keys; also raises the reference's EOF / identity
//                         conditions (multiexp.rs:55-65,74-80)
//   2. scan               exclusive prefix sums -> bucket offsets and task offsets
//   3. msm_scatter_kernel counting-sort scatter of (base index, sign) by bucket
//   4. msm_accumulate     one thread per <= L-point slice of a bucket: XYZZ += affine
//                         (8M + 2S), vectorised 16 B loads of the Montgomery coordinates
//   5. msm_combine_heavy  buckets split over several slices (skewed scalars) are tree-summed
//   6. msm_reduce_kernel  per window: sum_k k * B_k by running sums over bucket segments
//   7. msm_final_kernel   Horner over windows, one inversion -> canonical affine
//
// Results are independent of the window size (SURVEY 8a'/7), so c is tuned for the GPU and
// differs from the reference's ceil(ln n).
//
// Algorithmic work per dense point: 128 B (G1: 96 B base + 32 B scalar) /
// 224 B (G2); 48 000 MAC32 (G1) / 144 000 (G2) at 16 windows.
#pragma once
#include "encode.cuh"
#include "quad.cuh"

namespace bmpc {

__device__ __forceinline__ void load_scalar(const uint32_t* p, uint32_t s[8]) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w;
    s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w;
}

// ------------------------------------------------------------------------ point loads
template <class F> struct PointIO;
template <> struct PointIO<Fp> {
    static constexpr int WORDS = 24;  // 96 B
};
template <> struct PointIO<Fp2> {
    static constexpr int WORDS = 48;  // 192 B
};

template <class F>
__device__ __forceinline__ Affine<F> load_affine(const Affine<F>* p) {
    Affine<F> r;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint32_t* w = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int j = 0; j < PointIO<F>::WORDS / 4; j++) {
        uint4 v = __ldg(q + j);
        w[4 * j] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
    }
    return r;
}
// --------------------------------------------------------------------- 4. accumulate
// One thread per task; a task is a slice of <= L consecutive entries of one bucket in `sorted`,
// described by desc[t] (built by task_desc_kernel, big tasks first so warps stay balanced).
template <class F, bool COMPACT>
__global__ void __launch_bounds__(128)
msm_accumulate_kernel(const Affine<F>* bases, const uint32_t* sorted, const uint4* desc,
                      const uint32_t* ntasks_p, XYZZ<F>* partials) {
    uint32_t ntasks = *ntasks_p;
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntasks) return;
    uint4 d = __ldg(desc + t);     // {first entry, length, partial slot, bucket}; big tasks first
    uint32_t start = d.x, end = d.x + d.y;
    XYZZ<F> acc = XYZZ<F>::identity();
    for (uint32_t j = start; j < end; j++) {
        uint32_t e = __ldg(sorted + j);
        Affine<F> p = load_affine<F>(bases + (e & 0x7fffffffu));
        if (e & 0x80000000u) p.y = p.y.neg();
        acc.template add_affine_impl<COMPACT>(p);
    }
    store_struct(partials + d.z, acc);
}

// ------------------------------------------------------------------ 5. heavy buckets
// One block per heavy bucket (grid-stride over the list): partials[toff[b]] <- sum of the
// bucket's partial sums.  Typical heavy buckets hold 2..30 partials (the over-full low buckets of
// the top window); 0/1-heavy witnesses produce a few buckets with thousands, which the 128
// threads fold in strides before the shared-memory tree.
#define BMPC_HEAVY_WARP_MAX 64u
template <class F>
__global__ void __launch_bounds__(128)
msm_combine_heavy_kernel(const uint32_t* toff, const uint32_t* heavy_list,
                         const uint32_t* heavy_count, XYZZ<F>* partials) {
    extern __shared__ uint4 heavy_smem[];
    XYZZ<F>* sm = reinterpret_cast<XYZZ<F>*>(heavy_smem);
    const uint32_t nh = *heavy_count;
    // Pass 1, one WARP per bucket: buckets with a few dozen partial sums (with window tables the low
    // 2^12 buckets, which the short top window fills a second time: ~21 slices each at 2^24).  A
    // 128-thread block per bucket left three of its four warps idle in the tree: 1.2 ms for 4096
    // buckets.  Buckets with more than BMPC_HEAVY_WARP_MAX partial sums wait for pass 2.
    {
        const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
        XYZZ<F>* wsm = sm + 32u * wib;
        const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
        for (uint32_t h = blockIdx.x * (blockDim.x >> 5) + wib; h < nh; h += nwarps) {
            uint32_t b = heavy_list[h];
            uint32_t t0 = toff[b], t1 = toff[b + 1];
            uint32_t k = t1 - t0;
            if (k > BMPC_HEAVY_WARP_MAX) continue;
            uint32_t width = 1;                  // power of two >= min(k, 32)
            while (width < k && width < 32u) width <<= 1;
            if (lane < width) {
                XYZZ<F> acc = XYZZ<F>::identity();
                for (uint32_t t = t0 + lane; t < t1; t += 32u) {
                    XYZZ<F> v = load_struct(partials + t);
                    acc.add(v);
                }
                wsm[lane] = acc;
            }
            __syncwarp();
            for (uint32_t st = width >> 1; st > 0; st >>= 1) {
                if (lane < st) {
                    XYZZ<F> x = wsm[lane], y = wsm[lane + st];
                    x.add(y);
                    wsm[lane] = x;
                }
                __syncwarp();
            }
            if (lane == 0) store_struct(partials + t0, wsm[0]);
            __syncwarp();
        }
    }
    __syncthreads();
    // Pass 2, one BLOCK per bucket: the few buckets with hundreds or thousands of partial sums.
    for (uint32_t h = blockIdx.x; h < nh; h += gridDim.x) {
        uint32_t b = heavy_list[h];
        uint32_t t0 = toff[b], t1 = toff[b + 1];
        uint32_t k = t1 - t0;
        if (k <= BMPC_HEAVY_WARP_MAX) continue;   // block-uniform: done in pass 1
        uint32_t width = 1;                      // power of two >= min(k, 128)
        while (width < k && width < 128u) width <<= 1;
        if (threadIdx.x < width) {
            XYZZ<F> acc = XYZZ<F>::identity();
            for (uint32_t t = t0 + threadIdx.x; t < t1; t += 128u) {
                XYZZ<F> v = load_struct(partials + t);
                acc.add(v);
            }
            sm[threadIdx.x] = acc;
        }
        __syncthreads();
        for (uint32_t st = width >> 1; st > 0; st >>= 1) {
            if (threadIdx.x < st) {
                XYZZ<F> x = sm[threadIdx.x], y = sm[threadIdx.x + st];
                x.add(y);
                sm[threadIdx.x] = x;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) store_struct(partials + t0, sm[0]);
        __syncthreads();
    }
}

// ------------------------------------------------------------------- 6. window reduce
// Sum of a bucket's partial sums: few -> folded here; many -> the heavy combine left the total
// in the first slot.
template <class F>
__device__ __forceinline__ XYZZ<F> bucket_value(const XYZZ<F>* partials, const uint32_t* toff, uint32_t bidx) {
    uint32_t t0 = __ldg(toff + bidx), t1 = __ldg(toff + bidx + 1);
    XYZZ<F> v = XYZZ<F>::identity();
    if (t1 > t0) {
        v = load_struct(partials + t0);
        uint32_t k = (t1 - t0 <= BMPC_INLINE_PARTIALS) ? (t1 - t0) : 1u;
        for (uint32_t q = 1; q < k; q++) {
            XYZZ<F> u = load_struct(partials + t0 + q);
            v.add(u);
        }
    }
    return v;
}

// 6. Weighted sum of one bucket set: sum_k (k + 1) B_k, with no scalar multiplication anywhere.
// Thread t owns S = 2^s consecutive buckets [t S, (t+1) S): running sums give
//   run_t = sum B,   acc_t = sum (local index + 1) B;
// the thread's buckets really weigh t S more each, and  sum_t t S run_t = S sum_{t>=1} Suf_t  with
// Suf_t = sum_{i>=t} run_i: an inclusive suffix scan over the block (log2 steps in shared memory),
// s doublings, and the usual tree.  The block emits V = its weighted sum with block-local weights and
// R = the plain sum of its buckets; the final kernel repeats the same step over the blocks.
// (The previous version multiplied run_t by t S with a double-and-add ladder per thread: ~300 cold
// field products against 28 per bucket -- 2.2 ms for 2^18 buckets.)
// grid = (nblk, sets), blockDim = rblock (power of two <= 256), smem = rblock * sizeof(XYZZ).
template <class F>
__global__ void __launch_bounds__(256)
msm_reduce_kernel(const XYZZ<F>* partials, const uint32_t* toff, uint32_t B, uint32_t s_log,
                  XYZZ<F>* blk_V, XYZZ<F>* blk_R) {
    extern __shared__ uint4 reduce_smem[];
    XYZZ<F>* sm = reinterpret_cast<XYZZ<F>*>(reduce_smem);
    const uint32_t w = blockIdx.y, tid = threadIdx.x, nth = blockDim.x;
    const uint32_t S = 1u << s_log;
    const uint32_t lo = (blockIdx.x * nth + tid) * S;
    XYZZ<F> run = XYZZ<F>::identity(), acc = XYZZ<F>::identity();
    // The batched-affine accumulate kernel leaves AFFINE bucket sums (ZZ = ZZZ = 1): `run += v` is
    // then a mixed addition (10 products instead of 14).  Decided per warp so that the lanes never
    // split over two addition bodies (XYZZ accumulate kernel, folded heavy buckets: general path).
    const unsigned wmask = __activemask();
    const F one = F::one();
    for (uint32_t j = S; j > 0; j--) {      // local weight j
        XYZZ<F> v = XYZZ<F>::identity();
        if (lo < B) v = bucket_value<F>(partials, toff, w * B + lo + j - 1u);
        const bool ident = v.is_identity();
        const bool aff = ident || (v.ZZ == one && v.ZZZ == one);
        if (__all_sync(wmask, aff)) {
            Affine<F> a = ident ? Affine<F>::identity() : Affine<F>{v.X, v.Y};
            run.add_affine_cold(a);
        } else {
            run.add(v);
        }
        acc.add(run);
    }
    // inclusive suffix scan of run over the block
    sm[tid] = run;
    __syncthreads();
    for (uint32_t d = 1; d < nth; d <<= 1) {
        const bool has = tid + d < nth;
        XYZZ<F> t = XYZZ<F>::identity();
        if (has) t = sm[tid + d];
        __syncthreads();
        if (has) {
            run.add(t);
            sm[tid] = run;
        }
        __syncthreads();
    }
    if (tid == 0) store_struct(blk_R + (size_t)w * gridDim.x + blockIdx.x, run);
    if (tid >= 1) {
        for (uint32_t q = 0; q < s_log; q++) run = run.dbl();
        acc.add(run);
    }
    __syncthreads();
    sm[tid] = acc;
    __syncthreads();
    for (uint32_t st = nth >> 1; st > 0; st >>= 1) {
        if (tid < st) {
            XYZZ<F> x = sm[tid], y = sm[tid + st];
            x.add(y);
            sm[tid] = x;
        }
        __syncthreads();
    }
    if (tid == 0) store_struct(blk_V + (size_t)w * gridDim.x + blockIdx.x, sm[0]);
}

// ---------------------------------------------------------------- 6b. fold of block results
// The reduce kernel's suffix-scan step once more, over groups of `cnt` (power of two <= 256)
// consecutive block results of one set: element e of a group weighs e * 2^m_log more than its
// block-local weights say.  Emits one (V, R) pair per group; the final kernel then folds the groups
// (its m_log grows by log2 cnt).  Why a separate launch: one 256-thread block folding 256 results
// keeps 8 warps on ONE SM's multiplier pipe (0.7 ms measured); groups of 32 run as single warps on
// different SMs, where a product costs its latency only, and let the reduce kernel itself use
// small blocks (its scan levels, with every resident warp busy, are the expensive ones).
// grid = (groups, sets), blockDim = cnt, smem = cnt * sizeof(XYZZ).
template <class F>
__global__ void __launch_bounds__(256)
msm_fold_kernel(const XYZZ<F>* in_V, const XYZZ<F>* in_R, uint32_t n_in, uint32_t m_log,
                XYZZ<F>* out_V, XYZZ<F>* out_R) {
    extern __shared__ uint4 fold_smem[];
    XYZZ<F>* sm = reinterpret_cast<XYZZ<F>*>(fold_smem);
    const uint32_t h = blockIdx.y, tid = threadIdx.x, cnt = blockDim.x;
    const size_t e = (size_t)h * n_in + (size_t)blockIdx.x * cnt + tid;
    XYZZ<F> run = load_struct(in_R + e), acc = load_struct(in_V + e);
    sm[tid] = run;
    __syncthreads();
    for (uint32_t d = 1; d < cnt; d <<= 1) {
        const bool has = tid + d < cnt;
        XYZZ<F> t = XYZZ<F>::identity();
        if (has) t = sm[tid + d];
        __syncthreads();
        if (has) {
            run.add(t);
            sm[tid] = run;
        }
        __syncthreads();
    }
    if (tid == 0) store_struct(out_R + (size_t)h * gridDim.x + blockIdx.x, run);
    if (tid >= 1) {
        for (uint32_t q = 0; q < m_log; q++) run = run.dbl();
        acc.add(run);
    }
    __syncthreads();
    sm[tid] = acc;
    __syncthreads();
    for (uint32_t st = cnt >> 1; st > 0; st >>= 1) {
        if (tid < st) {
            XYZZ<F> x = sm[tid], y = sm[tid + st];
            x.add(y);
            sm[tid] = x;
        }
        __syncthreads();
    }
    if (tid == 0) store_struct(out_V + (size_t)h * gridDim.x + blockIdx.x, sm[0]);
}

// ------------------------------------------------------------------------ 7. final
// One block per bucket set: the same suffix-scan step over the set's nblk (<= 256, power of two)
// block results -- block b's buckets weigh b * 2^m_log more (m_log = log2(rblock S)) -- then the
// LAST block to finish (ticket) runs Horner over the sets (the doubling fold of
// multiexp.rs:244-249: c doublings between sets; with precomputed tables H == 1, none).
// mode 0: canonical affine, uncompressed big-endian bytes to out_bytes
// mode 1: leave the XYZZ partial in out_xyzz (sharded MSM)
// Launch: H blocks of FINAL_THREADS threads, dynamic smem = FINAL_THREADS * sizeof(XYZZ);
// `win` = H XYZZ slots in global memory, *ticket = 0 on entry.
#define BMPC_FINAL_THREADS 256
template <class F>
__global__ void __launch_bounds__(BMPC_FINAL_THREADS)
msm_final_kernel(const XYZZ<F>* blk_V, const XYZZ<F>* blk_R, uint32_t H, uint32_t nblk, uint32_t m_log,
                 uint32_t c, int mode, XYZZ<F>* win, uint32_t* ticket, uint8_t* out_bytes,
                 XYZZ<F>* out_xyzz) {
    extern __shared__ uint4 final_smem[];
    XYZZ<F>* sm = reinterpret_cast<XYZZ<F>*>(final_smem);
    __shared__ uint32_t last;
    const uint32_t h = blockIdx.x, tid = threadIdx.x;
    XYZZ<F> run = XYZZ<F>::identity(), acc = XYZZ<F>::identity();
    if (tid < nblk) {
        run = load_struct(blk_R + (size_t)h * nblk + tid);
        acc = load_struct(blk_V + (size_t)h * nblk + tid);
    }
    if (nblk > 1) {
        sm[tid] = run;
        __syncthreads();
        for (uint32_t d = 1; d < nblk; d <<= 1) {
            const bool has = tid + d < nblk;
            XYZZ<F> t = XYZZ<F>::identity();
            if (has) t = sm[tid + d];
            __syncthreads();
            if (has) {
                run.add(t);
                sm[tid] = run;
            }
            __syncthreads();
        }
        if (tid >= 1 && tid < nblk) {
            for (uint32_t q = 0; q < m_log; q++) run = run.dbl();
            acc.add(run);
        }
        __syncthreads();
        sm[tid] = acc;
        __syncthreads();
        for (uint32_t st = nblk >> 1; st > 0; st >>= 1) {
            if (tid < st) {
                XYZZ<F> x = sm[tid], y = sm[tid + st];
                x.add(y);
                sm[tid] = x;
            }
            __syncthreads();
        }
        if (tid == 0) acc = sm[0];
    }
    if (tid == 0) {
        store_struct(win + h, acc);
        __threadfence();
        last = (atomicAdd(ticket, 1u) == H - 1u);
    }
    __syncthreads();
    if (!last || tid != 0) return;
    __threadfence();
    XYZZ<F> tot = XYZZ<F>::identity();
    for (int hh = (int)H - 1; hh >= 0; hh--) {     // Horner from the top set down
        if (hh != (int)H - 1)
            for (uint32_t j = 0; j < c; j++) tot = tot.dbl();
        XYZZ<F> v = load_struct(win + hh);
        tot.add(v);
    }
    if (mode == 1) { store_struct(out_xyzz, tot); return; }
    Affine<F> a = tot.to_affine();
    encode_uncompressed<F>(a, out_bytes);
}

// ------------------------------------------------- 6b / 7 with four lanes per element (quad.cuh)
// The fold and final steps are pure latency: ~12 dependent additions and m_log doublings on one
// warp per group.  With a quad of lanes per element an addition costs 5 product latencies instead of
// 14 and a doubling 3 instead of 9 (2^19 buckets, G1: fold 0.30 -> 0.13 ms, final 0.50 -> 0.2 ms).
// Same suffix-scan step as msm_fold_kernel; blockDim.x = 4 * cnt, cnt a power of two >= 8 (elements
// beyond n_valid are the identity), smem = cnt * sizeof(XYZZ).
template <class F>
__device__ __forceinline__ void quad_scan_step(XYZZ<F>* sm, XYZZ<F>& run, XYZZ<F>& acc, uint32_t e, uint32_t role,
                                               uint32_t cnt, uint32_t m_log) {
    if (role == 0) sm[e] = run;
    __syncthreads();
    for (uint32_t d = 1; d < cnt; d <<= 1) {
        const bool has = e + d < cnt;
        XYZZ<F> t = XYZZ<F>::identity();
        if (has) t = sm[e + d];
        __syncthreads();
        quad_add<F>(run, t);
        if (has && role == 0) sm[e] = run;
        __syncthreads();
    }
    // run = suffix sum from e on; element e weighs e 2^m_log more than its local weights say:
    // sum_e e run_e = sum_{e >= 1} Suf_e
    XYZZ<F> w = e >= 1 ? run : XYZZ<F>::identity();
    for (uint32_t q = 0; q < m_log; q++) w = quad_dbl<F>(w);
    quad_add<F>(acc, w);
    __syncthreads();
    if (role == 0) sm[e] = acc;
    __syncthreads();
    for (uint32_t st = cnt >> 1; st > 0; st >>= 1) {
        XYZZ<F> x = sm[e], y = XYZZ<F>::identity();
        if (e < st) y = sm[e + st];
        __syncthreads();
        quad_add<F>(x, y);
        if (e < st && role == 0) sm[e] = x;
        __syncthreads();
    }
    // sm[0] = the group's weighted sum; `run` of element 0 = its plain sum
}

template <class F>
__global__ void __launch_bounds__(128)
msm_fold_quad_kernel(const XYZZ<F>* in_V, const XYZZ<F>* in_R, uint32_t n_in, uint32_t m_log,
                     XYZZ<F>* out_V, XYZZ<F>* out_R) {
    extern __shared__ uint4 foldq_smem[];
    XYZZ<F>* sm = reinterpret_cast<XYZZ<F>*>(foldq_smem);
    const uint32_t h = blockIdx.y, e = threadIdx.x >> 2, role = threadIdx.x & 3u, cnt = blockDim.x >> 2;
    const size_t idx = (size_t)h * n_in + (size_t)blockIdx.x * cnt + e;
    XYZZ<F> run = load_struct(in_R + idx), acc = load_struct(in_V + idx);
    quad_scan_step<F>(sm, run, acc, e, role, cnt, m_log);
    if (threadIdx.x == 0) {
        store_struct(out_R + (size_t)h * gridDim.x + blockIdx.x, run);
        store_struct(out_V + (size_t)h * gridDim.x + blockIdx.x, sm[0]);
    }
}

// final over n_valid <= 64 (V, R) pairs per set; blockDim.x = 4 * cnt with cnt = max(8, n_valid)
// rounded to a power of two.  The last block to finish runs Horner over the sets on its first quad.
template <class F>
__global__ void __launch_bounds__(256)
msm_final_quad_kernel(const XYZZ<F>* blk_V, const XYZZ<F>* blk_R, uint32_t H, uint32_t n_valid, uint32_t m_log,
                      uint32_t c, int mode, XYZZ<F>* win, uint32_t* ticket, uint8_t* out_bytes,
                      XYZZ<F>* out_xyzz) {
    extern __shared__ uint4 finalq_smem[];
    XYZZ<F>* sm = reinterpret_cast<XYZZ<F>*>(finalq_smem);
    __shared__ uint32_t last;
    const uint32_t h = blockIdx.x, e = threadIdx.x >> 2, role = threadIdx.x & 3u, cnt = blockDim.x >> 2;
    XYZZ<F> run = XYZZ<F>::identity(), acc = XYZZ<F>::identity();
    if (e < n_valid) {
        run = load_struct(blk_R + (size_t)h * n_valid + e);
        acc = load_struct(blk_V + (size_t)h * n_valid + e);
    }
    quad_scan_step<F>(sm, run, acc, e, role, cnt, m_log);
    if (threadIdx.x == 0) {
        store_struct(win + h, sm[0]);
        __threadfence();
        last = (atomicAdd(ticket, 1u) == H - 1u);
    }
    __syncthreads();
    if (!last || e != 0) return;
    __threadfence();
    XYZZ<F> tot = XYZZ<F>::identity();
    for (int hh = (int)H - 1; hh >= 0; hh--) {     // Horner from the top set down, on the first quad
        if (hh != (int)H - 1)
            for (uint32_t j = 0; j < c; j++) tot = quad_dbl<F>(tot);
        XYZZ<F> v = load_struct(win + hh);
        quad_add<F>(tot, v);
    }
    if (role != 0) return;
    if (mode == 1) { store_struct(out_xyzz, tot); return; }
    Affine<F> a = tot.to_affine();
    encode_uncompressed<F>(a, out_bytes);
}

// Precomputed window tables: tables[w][i] = 2^(c w) * P_i in affine form, w = 1 .. W-1 (table 0
// = the bases themselves).  One thread per base: chain of c doublings per table kept in XYZZ,
// then ONE inversion per base for all tables (Montgomery's trick over the W-1 denominators).
#define BMPC_MAX_TABLES 32
template <class F>
__global__ void __launch_bounds__(64)
msm_precompute_kernel(Affine<F>* tables, size_t n, uint32_t c, uint32_t W) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine<F> p = load_struct(tables + i);
    if (p.is_identity()) {
        for (uint32_t w = 1; w < W; w++) store_struct(tables + (size_t)w * n + i, p);
        return;
    }
    XYZZ<F> pts[BMPC_MAX_TABLES];
    F pre[BMPC_MAX_TABLES];
    XYZZ<F> acc = XYZZ<F>::from_affine(p);
    F prod = F::one();
    for (uint32_t w = 1; w < W; w++) {
        for (uint32_t j = 0; j < c; j++) acc = acc.dbl();
        pts[w] = acc;
        pre[w] = prod;                                  // product of the denominators before w
        if (!acc.is_identity()) prod = F::mul_cold(prod, F::mul_cold(acc.ZZ, acc.ZZZ));
    }
    F inv = prod.inv();
    for (uint32_t w = W - 1; w >= 1; w--) {
        Affine<F> out = Affine<F>::identity();
        if (!pts[w].is_identity()) {
            F d = F::mul_cold(pts[w].ZZ, pts[w].ZZZ);
            F di = F::mul_cold(inv, pre[w]);            // 1 / (ZZ * ZZZ) of table w
            inv = F::mul_cold(inv, d);
            out.x = F::mul_cold(pts[w].X, F::mul_cold(di, pts[w].ZZZ));
            out.y = F::mul_cold(pts[w].Y, F::mul_cold(di, pts[w].ZZ));
        }
        store_struct(tables + (size_t)w * n + i, out);
    }
}

// sum of `count` XYZZ partials (one per rank) -> canonical affine bytes.  One block of 32 threads:
// strided loads, shuffle-free shared-memory tree (depth 3 for 8 ranks instead of 8 serial additions).
template <class F>
__global__ void __launch_bounds__(32)
msm_sum_partials_kernel(const XYZZ<F>* parts, uint32_t count, uint8_t* out_bytes) {
    __shared__ uint4 sp_smem[32 * sizeof(XYZZ<F>) / 16];
    XYZZ<F>* sm = reinterpret_cast<XYZZ<F>*>(sp_smem);
    uint32_t width = 1;
    while (width < count && width < 32u) width <<= 1;
    if (threadIdx.x < width) {
        XYZZ<F> acc = XYZZ<F>::identity();
        for (uint32_t j = threadIdx.x; j < count; j += width) {
            XYZZ<F> v = load_struct(parts + j);
            acc.add(v);
        }
        sm[threadIdx.x] = acc;
    }
    __syncthreads();
    for (uint32_t st = width >> 1; st > 0; st >>= 1) {
        if (threadIdx.x < st) {
            XYZZ<F> x = sm[threadIdx.x], y = sm[threadIdx.x + st];
            x.add(y);
            sm[threadIdx.x] = x;
        }
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    Affine<F> a = sm[0].to_affine();
    encode_uncompressed<F>(a, out_bytes);
}

// ------------------------------------------------------------ base ingestion / export
// uncompressed big-endian -> Montgomery x|y ((0,0) for the infinity flag)
template <class F>
__global__ void decode_uncompressed_kernel(const uint8_t* in, size_t stride, size_t n,
                                           Affine<F>* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* p = in + i * stride;
    const int CB = sizeof(F);
    Affine<F> a;
    if (p[0] & 0x40) {
        a = Affine<F>::identity();
    } else {
        get_coord_be(p, 0x1f, a.x);
        get_coord_be(p + CB, 0xff, a.y);
    }
    store_struct(out + i, a);
}
// ---- validating decode (Parameters::read / VerifyingKey::read, groth16/mod.rs:161-221,292-400)
// G1Affine::from_uncompressed{,_unchecked} of bls12_381 0.6.0: flags, canonical coordinates,
// and (checked) on-curve + prime-order subgroup.  err <- min over failing points of
// index * 4 + kind, kind 1 = "invalid G1/G2", 2 = "point at infinity" (identity where rejected).
__device__ __forceinline__ bool fp_be_canonical(const uint8_t* in, uint8_t mask0) {
    // big-endian bytes < p ?
    bool less = false, decided = false;
    for (int j = 0; j < 12; j++) {
        uint32_t b0 = in[4 * j], b1 = in[4 * j + 1], b2 = in[4 * j + 2], b3 = in[4 * j + 3];
        if (j == 0) b0 &= mask0;
        uint32_t w = (b0 << 24) | (b1 << 16) | (b2 << 8) | b3;
        uint32_t m = FpParams::mod(11 - j);
        if (!decided && w != m) { less = w < m; decided = true; }
    }
    return decided && less;
}
__device__ __forceinline__ bool coord_canonical(const uint8_t* in, uint8_t mask0, const Fp&) {
    return fp_be_canonical(in, mask0);
}
__device__ __forceinline__ bool coord_canonical(const uint8_t* in, uint8_t mask0, const Fp2&) {
    return fp_be_canonical(in, mask0) && fp_be_canonical(in + 48, 0xff);
}
__device__ __forceinline__ Fp curve_b(const Fp&) {            // y^2 = x^3 + 4
    Fp one = Fp::one(), two = one + one;
    return two + two;
}
__device__ __forceinline__ Fp2 curve_b(const Fp2&) {          // y^2 = x^3 + 4 (1 + u)
    Fp four = curve_b(Fp::zero());
    return Fp2{four, four};
}
template <class F>
__global__ void __launch_bounds__(64)
validate_decode_kernel(const uint8_t* in, size_t stride, size_t n, int checked, int reject_identity,
                       Affine<F>* out, uint32_t* err) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* p = in + i * stride;
    const int CB = sizeof(F);
    uint32_t kind = 0;
    Affine<F> a = Affine<F>::identity();
    const uint32_t f = p[0];
    if (!coord_canonical(p, 0x1f, a.x) || !coord_canonical(p + CB, 0xff, a.y) || (f & 0x80) || (f & 0x20)) {
        kind = 1;
    } else {
        get_coord_be(p, 0x1f, a.x);
        get_coord_be(p + CB, 0xff, a.y);
        bool zero = a.x.is_zero() && a.y.is_zero();
        if (f & 0x40) {
            if (!zero) kind = 1;
            else if (reject_identity) kind = 2;
        } else if (zero) {
            kind = 1;   // (0,0) without the infinity flag: not a curve point (and our identity sentinel)
        } else if (checked) {
            F lhs = F::mul_cold(a.y, a.y);
            F rhs = F::mul_cold(F::mul_cold(a.x, a.x), a.x) + curve_b(a.x);
            if (lhs != rhs) kind = 1;
            else {
                uint32_t r[8];
                for (int j = 0; j < 8; j++) r[j] = FrParams::mod(j);
                XYZZ<F> t = XYZZ<F>::from_affine(a).mul(r, 8);     // [r] P == O  <=>  torsion free
                if (!t.is_identity()) kind = 1;
            }
        }
    }
    if (kind) atomicMin(err, (uint32_t)(i * 4 + kind));
    store_struct(out + i, kind ? Affine<F>::identity() : a);
}

template <class F>
__global__ void encode_uncompressed_kernel(const Affine<F>* in, size_t n, uint8_t* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine<F> a = load_struct(in + i);
    encode_uncompressed<F>(a, out + i * 2 * sizeof(F));
}
// bit i of inf_bitmap = bases[i] is the identity.  One warp per 32 bases (ballot).
template <class F>
__global__ void inf_bitmap_kernel(const Affine<F>* bases, size_t n, uint32_t* bitmap) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool inf = false;
    if (i < n) {
        Affine<F> a = load_struct(bases + i);
        inf = a.is_identity();
    }
    uint32_t m = __ballot_sync(0xffffffffu, inf);
    if ((threadIdx.x & 31) == 0 && i < n) bitmap[i >> 5] = m;
}

// -------------------------------------------------------------- batch scalar multiply
// out[i] = in[i] * k[i or 0]  (reference: mpc.rs:647-706 per-element `Mul`); canonical
// affine out (one inversion per point -- setup-time path).
template <class F>
__global__ void __launch_bounds__(128)
batch_scalar_mul_kernel(const Affine<F>* in, const uint32_t* scalars, int per_element, size_t n,
                        Affine<F>* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s[8];
    load_scalar(scalars + (per_element ? i * 8 : 0), s);
    Affine<F> p = load_struct(in + i);
    XYZZ<F> acc = XYZZ<F>::identity();
    if (!p.is_identity()) {
        int top = 255;
        while (top >= 0 && !((s[top >> 5] >> (top & 31)) & 1u)) top--;
        for (int b = top; b >= 0; b--) {
            acc = acc.dbl();
            if ((s[b >> 5] >> (b & 31)) & 1u) acc.add_affine(p);
        }
    }
    store_struct(out + i, acc.to_affine());
}

// ------------------------------------------------------------------ list x sparse matrix
// out[i] = sum_j coeff[j] * list[col[j]] over row i's entries j in [row_ptr[i], row_ptr[i+1])
// (reference: mpc.rs:416-457 list_mul_matrix -- a segmented multiexp whose segments are the rows of
// a QAP matrix, a handful of entries each).  One thread per row, joint double-and-add over the row
// (the 255 doublings are shared by all its entries); rows >= live_rows (the reference stops at the
// first empty row, :432-434) and the tail up to the list's length stay the identity.  Canonical
// affine out.
template <class F>
__global__ void __launch_bounds__(128)
list_mul_matrix_kernel(const Affine<F>* list, const uint32_t* row_ptr, const uint32_t* col, const uint32_t* coeffs,
                       size_t live_rows, size_t n_out, Affine<F>* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    XYZZ<F> acc = XYZZ<F>::identity();
    if (i < live_rows) {
        const uint32_t j0 = row_ptr[i], j1 = row_ptr[i + 1];
        int top = -1;                                   // highest set bit over the row's coefficients
        for (uint32_t j = j0; j < j1; j++) {
            if (load_struct(list + col[j]).is_identity()) continue;
            for (int w = 7; w >= 0 && w * 32 + 31 > top; w--) {
                uint32_t v = coeffs[(size_t)j * 8 + w];
                if (v) { int t = w * 32 + 31 - __clz(v); if (t > top) top = t; break; }
            }
        }
        for (int b = top; b >= 0; b--) {
            acc = acc.dbl();
            for (uint32_t j = j0; j < j1; j++) {
                if (!((coeffs[(size_t)j * 8 + (b >> 5)] >> (b & 31)) & 1u)) continue;
                Affine<F> p = load_struct(list + col[j]);
                if (!p.is_identity()) acc.add_affine_cold(p);
            }
        }
    }
    store_struct(out + i, acc.to_affine());
}

// Fixed-base: table[w][d] = base * (d * 2^(8 w)), d in [0,256), w in [0,32) as XYZZ.
template <class F>
__global__ void fixed_base_table_kernel(const Affine<F>* base, XYZZ<F>* table) {
    // one thread per window: walks d = 1..255 by repeated addition of base * 2^(8w)
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= 32) return;
    XYZZ<F> step = XYZZ<F>::from_affine(*base);
    for (uint32_t j = 0; j < 8 * w; j++) step = step.dbl();
    XYZZ<F> cur = XYZZ<F>::identity();
    for (uint32_t d = 0; d < 256; d++) {
        store_struct(table + (size_t)w * 256 + d, cur);
        cur.add(step);
    }
}
template <class F>
__global__ void __launch_bounds__(128)
fixed_base_mul_kernel(const XYZZ<F>* table, const uint32_t* scalars, size_t n, Affine<F>* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s[8];
    load_scalar(scalars + i * 8, s);
    XYZZ<F> acc = XYZZ<F>::identity();
    for (uint32_t w = 0; w < 32; w++) {
        uint32_t d = (s[w >> 2] >> ((w & 3) * 8)) & 0xffu;
        if (d) {
            XYZZ<F> v = load_struct(table + (size_t)w * 256 + d);
            acc.add(v);
        }
    }
    store_struct(out + i, acc.to_affine());
}

}  // namespace bmpc
