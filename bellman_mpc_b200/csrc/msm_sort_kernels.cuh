// Digit extraction, density -> base index resolution and counting sort for the MSM
// (stages 1-3 of the pipeline described in group_kernels.cuh).
#pragma once
#include "internal.h"
#include "msm_pair_lists.h"

namespace bmpc {

struct MsmInput {
    const uint32_t* scalars;       // n x 8 u32 canonical little-endian
    size_t n;
    const uint32_t* density;       // NULL = FullDensity; else 32-bit view of the u64 words
    const uint32_t* word_prefix;   // exclusive popcount prefix per 32-bit density word
    uint32_t base_offset;
    uint32_t bases_len;
    const uint32_t* inf_bitmap;    // bit per base: 1 = identity
};

__device__ __forceinline__ void load_scalar(const uint32_t* p, uint32_t s[8]) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w;
    s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w;
}

// bits [pos, pos + c) of a 256-bit little-endian integer (bits >= 256 read as zero), c <= 24
__device__ __forceinline__ uint32_t get_bits(const uint32_t s[8], uint32_t pos, uint32_t c) {
    uint32_t w = pos >> 5, sh = pos & 31;
    uint32_t lo = w < 8 ? s[w] : 0u;
    uint32_t hi = (w + 1) < 8 ? s[w + 1] : 0u;
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (uint32_t)(v >> sh) & ((1u << c) - 1u);
}

// Resolves position i -> (dense?, base index); returns false if nothing is consumed.
__device__ __forceinline__ bool resolve_base(const MsmInput& in, size_t i, uint32_t& base_idx) {
    if (in.density) {
        uint32_t wi = (uint32_t)(i >> 5), bit = (uint32_t)(i & 31);
        uint32_t word = __ldg(in.density + wi);
        if (!((word >> bit) & 1u)) return false;
        uint32_t rank = __ldg(in.word_prefix + wi) + __popc(word & ((1u << bit) - 1u));
        base_idx = in.base_offset + rank;
    } else {
        base_idx = in.base_offset + (uint32_t)i;
    }
    return true;
}

// Shared prologue of count and scatter: resolves the base, applies the reference's skip /
// EOF / identity rules and loads the scalar.  Returns false when the position contributes
// nothing (the caller keeps running with `alive == false`: both kernels use full-warp
// match_any, so no lane may leave early).
__device__ __forceinline__ bool msm_prepare(const MsmInput& in, const MsmGeom& g, size_t i,
                                            uint32_t* flags, bool raise, uint32_t& base_idx,
                                            uint32_t s[8]) {
    if (i >= in.n) return false;
    if (!resolve_base(in, i, base_idx)) return false;
    if (base_idx >= in.bases_len) {  // Source::{next,skip} -> UnexpectedEof (multiexp.rs:55-61,74-80)
        if (raise) atomicOr(flags, (uint32_t)MSM_FLAG_EOF);
        return false;
    }
    load_scalar(in.scalars + i * 8, s);
    if ((s[0] | s[1] | s[2] | s[3] | s[4] | s[5] | s[6] | s[7]) == 0) return false;  // skip(1)
    if ((__ldg(in.inf_bitmap + (base_idx >> 5)) >> (base_idx & 31)) & 1u) {
        // next() on an identity base -> UnexpectedIdentity (multiexp.rs:63-65); which error wins
        // depends on whether the reference's top window would have consumed it.
        if (raise) {
            uint32_t fl = MSM_FLAG_IDENT_ANY;
            if (get_bits(s, g.top_skip, g.c_ref) != 0) fl |= MSM_FLAG_IDENT_TOP;
            atomicOr(flags, fl);
        }
        return false;
    }
    return true;
}

// Warp-aggregated atomicAdd: lanes of the warp holding the same key are grouped with
// match.any, one lane adds the group size and every lane gets base + its rank in the group.
// With uniform scalars groups are singletons; with 0/1-heavy witnesses (most lanes hit the same
// bucket) it removes the same-address serialisation in L2.  Must be called by all 32 lanes;
// dead lanes pass live = false.
__device__ __forceinline__ uint32_t warp_agg_add(uint32_t* counters, uint32_t key, bool live, bool want_pos) {
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t k = live ? key : (0xffffffe0u + lane);       // unique dummy keys for dead lanes
    uint32_t grp = __match_any_sync(0xffffffffu, k);
    uint32_t leader = __ffs(grp) - 1u;
    uint32_t base = 0;
    if (live && lane == leader) base = atomicAdd(counters + key, (uint32_t)__popc(grp));
    if (!want_pos) return 0;
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + __popc(grp & ((1u << lane) - 1u));
}

// window w -> bucket set (w % H), table (w / H): with precomputed tables H == 1 and every window
// adds table_w[i] = 2^(c w) P_i into the same bucket set.
__global__ void __launch_bounds__(256)
msm_count_kernel(MsmInput in, MsmGeom g, uint32_t* hist, uint32_t* flags) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t base_idx = 0, s[8];
    bool alive = msm_prepare(in, g, i, flags, true, base_idx, s);
    uint32_t carry = 0;
    const uint32_t half = 1u << (g.c - 1);
    for (uint32_t w = 0; w < g.W; w++) {
        uint32_t d = 0;
        if (alive) {
            d = get_bits(s, w * g.c, g.c) + carry;
            if (d > half) { d = (1u << g.c) - d; carry = 1; } else carry = 0;
        }
        warp_agg_add(hist, (w % g.H) * g.B + (d - 1u), alive && d != 0, false);
    }
}

// Scatter: the slot of each (position, window) pair comes back from the (aggregated) atomicAdd.
__global__ void __launch_bounds__(256)
msm_scatter_kernel(MsmInput in, MsmGeom g, uint32_t* cursor, uint32_t* sorted) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t base_idx = 0, s[8];
    bool alive = msm_prepare(in, g, i, nullptr, false, base_idx, s);
    uint32_t carry = 0;
    const uint32_t half = 1u << (g.c - 1);
    for (uint32_t w = 0; w < g.W; w++) {
        uint32_t d = 0;
        bool neg = false;
        if (alive) {
            d = get_bits(s, w * g.c, g.c) + carry;
            neg = d > half;
            if (neg) { d = (1u << g.c) - d; carry = 1; } else carry = 0;
        }
        bool live = alive && d != 0;
        uint32_t pos = warp_agg_add(cursor, (w % g.H) * g.B + (d - 1u), live, true);
        if (live) sorted[pos] = (base_idx + (w / g.H) * g.tab_stride) | (neg ? 0x80000000u : 0u);
    }
}

// ---------------------------------------------------------------- density prefix popcount
__global__ void popc_words_kernel(const uint32_t* words, uint32_t nwords, uint32_t nbits,
                                  uint32_t* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nwords) return;
    uint32_t w = words[i];
    uint32_t rem = nbits - i * 32u;           // bits of this word that are inside the map
    if (rem < 32u) w &= (1u << rem) - 1u;
    out[i] = __popc(w);
}

// ------------------------------------------------------------------ exclusive scan (u32)
// Three phases over chunks of SCAN_CHUNK elements; `mode` 1 scans ceil(x / L) instead of x.
#define BMPC_SCAN_THREADS 256
#define BMPC_SCAN_ITEMS 4
#define BMPC_SCAN_CHUNK (BMPC_SCAN_THREADS * BMPC_SCAN_ITEMS)

#define BMPC_SCAN_EVEN 0xffffffffu       // scan x rounded up to even (pair mode bucket offsets)
__device__ __forceinline__ uint32_t scan_xform(uint32_t x, uint32_t L) {
    if (L == BMPC_SCAN_EVEN) return x + (x & 1u);
    return L ? (x + L - 1u) / L : x;
}

// returns exclusive prefix of `v` across the block; total in *total (valid in all threads)
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t warp_sums[BMPC_SCAN_THREADS / 32];
    __shared__ uint32_t block_total;
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t ws = lane < BMPC_SCAN_THREADS / 32 ? warp_sums[lane] : 0u;
        uint32_t winc = ws;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= (uint32_t)o) winc += t;
        }
        if (lane < BMPC_SCAN_THREADS / 32) warp_sums[lane] = winc - ws;
        if (lane == 31) block_total = winc;
    }
    __syncthreads();
    uint32_t r = inc - v + warp_sums[wid];
    *total = block_total;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(BMPC_SCAN_THREADS)
scan_phase1_kernel(const uint32_t* in, uint32_t n, uint32_t L, uint32_t* chunk_sums) {
    uint32_t base = blockIdx.x * BMPC_SCAN_CHUNK + threadIdx.x * BMPC_SCAN_ITEMS;
    uint32_t s = 0;
    for (int j = 0; j < BMPC_SCAN_ITEMS; j++)
        if (base + j < n) s += scan_xform(in[base + j], L);
    uint32_t total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) chunk_sums[blockIdx.x] = total;
}
// single block: exclusive scan of chunk_sums in place; grand total -> chunk_sums[nchunks]
__global__ void __launch_bounds__(BMPC_SCAN_THREADS)
scan_phase2_kernel(uint32_t* chunk_sums, uint32_t nchunks) {
    uint32_t carry = 0;
    for (uint32_t start = 0; start < nchunks; start += BMPC_SCAN_THREADS) {
        uint32_t idx = start + threadIdx.x;
        uint32_t v = idx < nchunks ? chunk_sums[idx] : 0u;
        uint32_t total;
        uint32_t ex = block_exclusive_scan(v, &total);
        if (idx < nchunks) chunk_sums[idx] = ex + carry;
        carry += total;
    }
    if (threadIdx.x == 0) chunk_sums[nchunks] = carry;
}
// out has n + 1 entries; out[n] = grand total
__global__ void __launch_bounds__(BMPC_SCAN_THREADS)
scan_phase3_kernel(const uint32_t* in, uint32_t n, uint32_t L, const uint32_t* chunk_sums,
                   uint32_t nchunks, uint32_t* out) {
    uint32_t base = blockIdx.x * BMPC_SCAN_CHUNK + threadIdx.x * BMPC_SCAN_ITEMS;
    uint32_t v[BMPC_SCAN_ITEMS];
    uint32_t s = 0;
    for (int j = 0; j < BMPC_SCAN_ITEMS; j++) {
        v[j] = (base + j < n) ? scan_xform(in[base + j], L) : 0u;
        s += v[j];
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan(s, &total) + chunk_sums[blockIdx.x];
    for (int j = 0; j < BMPC_SCAN_ITEMS; j++) {
        if (base + j < n) out[base + j] = ex;
        ex += v[j];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = chunk_sums[nchunks];
}

// ------------------------------------------------------------- task ordering by size
// Threads of a warp run until the longest of their 32 tasks finishes, so tasks are handed out
// in (approximately) descending size order: a 64-bin counting sort of the task lengths.  Each
// task gets a descriptor {first sorted entry, length, partial-sum slot}.
#define BMPC_TASK_BINS 64
__device__ __forceinline__ uint32_t task_bin(uint32_t len, uint32_t L) {
    return (BMPC_TASK_BINS - 1u) - (len * (BMPC_TASK_BINS - 1u)) / L;   // big tasks -> low bins
}
__global__ void __launch_bounds__(256)
task_bin_count_kernel(const uint32_t* off, const uint32_t* toff, uint32_t nb, uint32_t L, uint32_t* bin_hist) {
    __shared__ uint32_t sh[BMPC_TASK_BINS];
    if (threadIdx.x < BMPC_TASK_BINS) sh[threadIdx.x] = 0;
    __syncthreads();
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nb) {
        uint32_t cnt = off[b + 1] - off[b];
        uint32_t nt = toff[b + 1] - toff[b];
        if (nt > 1) atomicAdd(&sh[task_bin(L, L)], nt - 1u);
        if (nt > 0) atomicAdd(&sh[task_bin(cnt - (nt - 1u) * L, L)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < BMPC_TASK_BINS && sh[threadIdx.x]) atomicAdd(bin_hist + threadIdx.x, sh[threadIdx.x]);
}
// bin_hist[BINS] -> exclusive starts in place (single thread; 64 entries)
__global__ void task_bin_scan_kernel(uint32_t* bin_hist) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t acc = 0;
    for (int i = 0; i < BMPC_TASK_BINS; i++) {
        uint32_t v = bin_hist[i];
        bin_hist[i] = acc;
        acc += v;
    }
}
__global__ void __launch_bounds__(256)
task_desc_kernel(const uint32_t* off, const uint32_t* toff, uint32_t nb, uint32_t L, uint32_t* bin_cursor,
                 uint4* desc) {
    __shared__ uint32_t sh_cnt[BMPC_TASK_BINS];
    __shared__ uint32_t sh_base[BMPC_TASK_BINS];
    if (threadIdx.x < BMPC_TASK_BINS) sh_cnt[threadIdx.x] = 0;
    __syncthreads();
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t cnt = 0, nt = 0, o = 0, t0 = 0, my_full = 0, my_last = 0;
    if (b < nb) {
        o = off[b];
        cnt = off[b + 1] - o;
        t0 = toff[b];
        nt = toff[b + 1] - t0;
        if (nt > 1) my_full = atomicAdd(&sh_cnt[task_bin(L, L)], nt - 1u);
        if (nt > 0) my_last = atomicAdd(&sh_cnt[task_bin(cnt - (nt - 1u) * L, L)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < BMPC_TASK_BINS) {
        uint32_t c = sh_cnt[threadIdx.x];
        sh_base[threadIdx.x] = c ? atomicAdd(bin_cursor + threadIdx.x, c) : 0u;
    }
    __syncthreads();
    if (b < nb && nt > 0) {
        uint32_t fb = sh_base[task_bin(L, L)] + my_full;
        for (uint32_t k = 0; k + 1 < nt; k++) desc[fb + k] = make_uint4(o + k * L, L, t0 + k, b);
        uint32_t last_len = cnt - (nt - 1u) * L;
        desc[sh_base[task_bin(last_len, L)] + my_last] = make_uint4(o + (nt - 1u) * L, last_len, t0 + nt - 1u, b);
    }
}

// buckets whose points were split over more than BMPC_INLINE_PARTIALS accumulate tasks
__global__ void msm_find_heavy_kernel(const uint32_t* toff, uint32_t nb, uint32_t* heavy_list,
                                      uint32_t* heavy_count) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    // up to BMPC_INLINE_PARTIALS partial sums are folded by the reduce kernel itself
    if (toff[b + 1] - toff[b] > BMPC_INLINE_PARTIALS) heavy_list[atomicAdd(heavy_count, 1u)] = b;
}

// ============================================================== pair lists (msm_pairs.cuh)
// Round-based accumulation: per task (bucket slice, desc[t] = {first entry, padded length, slot,
// bucket}) the number of pairs in round r is pair_count(length, r); rows r = 1 .. R-1 are scanned
// over the tasks (blockIdx.y = r - 1) so that every round has ONE dense list.  `stride` = row pitch of
// pairoff / chunk_sums.

__global__ void __launch_bounds__(BMPC_SCAN_THREADS)
pair_scan_phase1_kernel(const uint4* desc, const uint32_t* ntasks_p, uint32_t* chunk_sums, uint32_t cstride) {
    const uint32_t n = *ntasks_p, r = blockIdx.y + 1;
    uint32_t base = blockIdx.x * BMPC_SCAN_CHUNK + threadIdx.x * BMPC_SCAN_ITEMS;
    uint32_t s = 0;
    for (int j = 0; j < BMPC_SCAN_ITEMS; j++)
        if (base + j < n) s += pair_count(__ldg(desc + base + j).y, r);
    uint32_t total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) chunk_sums[(size_t)blockIdx.y * cstride + blockIdx.x] = total;
}
// one block per row: exclusive scan of the row's chunk sums in place; row total -> totals[r]
__global__ void __launch_bounds__(BMPC_SCAN_THREADS)
pair_scan_phase2_kernel(uint32_t* chunk_sums, uint32_t cstride, uint32_t nchunks, uint32_t* totals) {
    uint32_t* row = chunk_sums + (size_t)blockIdx.x * cstride;
    uint32_t carry = 0;
    for (uint32_t start = 0; start < nchunks; start += BMPC_SCAN_THREADS) {
        uint32_t idx = start + threadIdx.x;
        uint32_t v = idx < nchunks ? row[idx] : 0u;
        uint32_t total;
        uint32_t ex = block_exclusive_scan(v, &total);
        if (idx < nchunks) row[idx] = ex + carry;
        carry += total;
    }
    if (threadIdx.x == 0) totals[blockIdx.x + 1] = carry;
}
__global__ void __launch_bounds__(BMPC_SCAN_THREADS)
pair_scan_phase3_kernel(const uint4* desc, const uint32_t* ntasks_p, const uint32_t* chunk_sums, uint32_t cstride,
                        uint32_t* pairoff, uint32_t stride) {
    const uint32_t n = *ntasks_p, r = blockIdx.y + 1;
    uint32_t base = blockIdx.x * BMPC_SCAN_CHUNK + threadIdx.x * BMPC_SCAN_ITEMS;
    uint32_t v[BMPC_SCAN_ITEMS];
    uint32_t s = 0;
    for (int j = 0; j < BMPC_SCAN_ITEMS; j++) {
        v[j] = (base + j < n) ? pair_count(__ldg(desc + base + j).y, r) : 0u;
        s += v[j];
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan(s, &total) + chunk_sums[(size_t)blockIdx.y * cstride + blockIdx.x];
    for (int j = 0; j < BMPC_SCAN_ITEMS; j++) {
        if (base + j < n) pairoff[(size_t)blockIdx.y * stride + base + j] = ex;
        ex += v[j];
    }
}

struct PairLayout {                       // by value into pair_build_kernel
    uint32_t R;
    uint32_t out_base[BMPC_PAIR_MAX_ROUNDS + 1];   // pool index of round r's first result
    uint32_t list_off[BMPC_PAIR_MAX_ROUNDS + 1];   // first entry of round r's list in `lists` (r >= 1)
};
// one thread per task: the lists of rounds 1 .. R-1 and the pool index of the task's sum
__global__ void __launch_bounds__(128)
pair_build_kernel(const uint4* desc, const uint32_t* ntasks_p, PairLayout lay, const uint32_t* pairoff,
                  uint32_t stride, uint2* lists, uint32_t* fin) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= *ntasks_p) return;
    const uint4 d = __ldg(desc + t);
    uint32_t po[BMPC_PAIR_MAX_ROUNDS + 1];
    PairIdx* lp[BMPC_PAIR_MAX_ROUNDS + 1];
    po[0] = 0;
    lp[0] = nullptr;
    for (uint32_t r = 1; r < lay.R; r++) {
        po[r] = pairoff[(size_t)(r - 1) * stride + t];
        lp[r] = reinterpret_cast<PairIdx*>(lists + lay.list_off[r]);
    }
    fin[t] = pair_build_task(d.x, d.y, lay.R, po, lay.out_base, lp);
}

}  // namespace bmpc
