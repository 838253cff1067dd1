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

// Warp-aggregated atomicAdd for skewed keys.  `match.any` groups ALL equal keys but costs a step per
// distinct value in the warp (uniform scalars: 32 distinct buckets, ~250 cycles per call -- it was the
// largest item of the count kernel and of every rs_* kernel when they used it).  What has to be
// defused is one DOMINANT key (0/1-heavy witnesses: most lanes hit bucket 0; the short top window: all
// lanes hit the same bin): the lanes holding the first live lane's key are found with one ballot and
// added by their leader, every other lane adds for itself.  Must be called by all 32 lanes; dead
// lanes pass live = false.
__device__ __forceinline__ uint32_t warp_agg_add(uint32_t* counters, uint32_t key, bool live, bool want_pos) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lv = __ballot_sync(0xffffffffu, live);
    if (!lv) return 0;
    const uint32_t src = __ffs(lv) - 1u;
    const uint32_t k0 = __shfl_sync(0xffffffffu, key, src);
    const uint32_t grp = __ballot_sync(0xffffffffu, live && key == k0);
    const bool in_grp = (grp >> lane) & 1u;
    uint32_t base = 0;
    if (lane == src) base = atomicAdd(counters + k0, (uint32_t)__popc(grp));
    else if (live && !in_grp) base = atomicAdd(counters + key, 1u);
    if (!want_pos) return 0;
    uint32_t gbase = __shfl_sync(0xffffffffu, base, src);
    return in_grp ? gbase + __popc(grp & ((1u << lane) - 1u)) : base;
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

// returns exclusive prefix of `v` across the block of NT threads; total in *total (valid in all threads)
template <int NT = BMPC_SCAN_THREADS>
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t warp_sums[NT / 32];
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
        uint32_t ws = lane < NT / 32 ? warp_sums[lane] : 0u;
        uint32_t winc = ws;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= (uint32_t)o) winc += t;
        }
        if (lane < NT / 32) warp_sums[lane] = winc - ws;
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

// ================================================================ two-level partition sort
// The one-pass scatter above writes every (position, window) pair to a random 4-byte slot of an
// array far larger than L2 (2^24 points: 805 MB): each store dirties a 32-byte sector that is
// evicted long before its other seven words arrive, so DRAM sees a read-modify-write per pair, and
// every pair pays a returning atomic on a random cursor (6.3 ms at 2^24, long_scoreboard 142 cycles
// per issue).  Here the pairs are first partitioned by BIN (RS_BIN_BUCKETS consecutive buckets) with
// tile-local staging in shared memory -- runs of a bin leave the SM as contiguous 8-byte entries
// {payload, bucket} -- and then scattered bin by bin: a block takes a chunk of one bin's entries,
// ranks them on shared-memory cursors and claims its output ranges with ONE global atomic per
// non-empty bucket of the chunk; its stores fall into the bin's slice of `sorted` (~400 KB at 2^24
// points), and what all resident blocks have in flight stays in L2 until the bin's other chunks have
// completed the sectors (see the chunk size note in msm_sort.cu).  The per-bucket histogram (bucket offsets, task
// offsets) comes from the same chunks instead of 2^24 x W global atomics.
//
//   rs_bin_count    positions -> bin histogram (+ the EOF / identity flags)
//   rs_bin_scan     bin offsets, bin cursors
//   rs_partition    positions -> entries grouped by bin
//   rs_bucket_hist  entries  -> per-bucket counts
//   (scan: off, toff as before)
//   rs_scatter      entries  -> sorted
#define RS_BIN_LOG 10u
#define RS_BIN_BUCKETS (1u << RS_BIN_LOG)
#define RS_MAX_BINS 2048u
#define RS_THREADS 256u
#define RS_PART_THREADS 512u   // rs_partition: 16 warps per block, two blocks per SM
#define RS_PPT4_MAX_WINDOWS 13u   // tile of 1024 positions: 1024 x 13 entries = 104 KB of staging, two blocks per SM
#define RS_MAX_WINDOWS 27u        // tile of 512 positions: 512 x 27 entries = 108 KB

// Shared-memory counters take plain atomics: lanes hitting the same counter are serialised by the
// hardware at a few cycles each, which even for a warp full of one key costs no more than the ballots
// and shuffles of an aggregated add -- and for the common case (distinct keys) a single instruction
// without a warp-wide dependency, so the walks of a thread's positions overlap.
__device__ __forceinline__ uint32_t warp_agg_add_sh(uint32_t* sh, uint32_t key, bool live, bool want_pos) {
    uint32_t pos = 0;
    if (live) pos = atomicAdd(sh + key, 1u);
    return pos;
}

// The signed digits of one scalar, window by window (the three position-side rs_* kernels).  The
// scalar is consumed by shifting a register copy right by c bits per window -- no indexed access to
// the limbs (local memory) and no division for the window's bucket set / table.  A dead position
// walks a zero scalar: every digit is 0.
struct DigitWalk {
    uint32_t r[8];
    uint32_t carry, set_base, tab_off;
    __device__ __forceinline__ void init(const uint32_t s[8], bool alive) {
#pragma unroll
        for (int k = 0; k < 8; k++) r[k] = alive ? s[k] : 0u;
        carry = 0;
        set_base = 0;
        tab_off = 0;
    }
    // digit d of the next window (0: nothing to add), its sign, its bucket key and its table offset
    __device__ __forceinline__ uint32_t next(const MsmGeom& g, bool& neg, uint32_t& key, uint32_t& tab) {
        uint32_t d = (r[0] & ((1u << g.c) - 1u)) + carry;
#pragma unroll
        for (int k = 0; k < 7; k++) r[k] = __funnelshift_r(r[k], r[k + 1], g.c);
        r[7] >>= g.c;
        neg = d > (g.B);                 // B = 2^(c-1)
        if (neg) { d = (1u << g.c) - d; carry = 1; } else carry = 0;
        key = set_base + (d - 1u);
        tab = tab_off;
        set_base += g.B;
        if (set_base == g.H * g.B) { set_base = 0; tab_off += g.tab_stride; }
        return d;
    }
};

template <int PPT>
__global__ void __launch_bounds__(RS_THREADS)
rs_bin_count_kernel(MsmInput in, MsmGeom g, uint32_t nbins, uint32_t* bin_hist, uint32_t* flags) {
    extern __shared__ uint32_t rs_sh[];
    uint32_t* hist = rs_sh;
    for (uint32_t b = threadIdx.x; b < nbins; b += RS_THREADS) hist[b] = 0;
    __syncthreads();
    const size_t tile0 = (size_t)blockIdx.x * (RS_THREADS * PPT);
    DigitWalk dw[PPT];
#pragma unroll
    for (int p = 0; p < PPT; p++) {
        size_t i = tile0 + (size_t)p * RS_THREADS + threadIdx.x;
        uint32_t base_idx = 0, s[8];
        bool alive = msm_prepare(in, g, i, flags, true, base_idx, s);
        dw[p].init(s, alive);
    }
    for (uint32_t w = 0; w < g.W; w++) {
#pragma unroll
        for (int p = 0; p < PPT; p++) {
            bool neg;
            uint32_t key, tab;
            uint32_t d = dw[p].next(g, neg, key, tab);
            warp_agg_add_sh(hist, key >> RS_BIN_LOG, d != 0, false);
        }
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nbins; b += RS_THREADS)
        if (hist[b]) atomicAdd(bin_hist + b, hist[b]);
}

// single block: bin_off[nbins + 1] = exclusive scan of bin_hist; bin_cursor = copy of the starts
__global__ void __launch_bounds__(BMPC_SCAN_THREADS)
rs_bin_scan_kernel(const uint32_t* bin_hist, uint32_t nbins, uint32_t* bin_off, uint32_t* bin_cursor) {
    uint32_t carry = 0;
    for (uint32_t start = 0; start < nbins; start += BMPC_SCAN_THREADS) {
        uint32_t idx = start + threadIdx.x;
        uint32_t v = idx < nbins ? bin_hist[idx] : 0u;
        uint32_t total;
        uint32_t ex = block_exclusive_scan(v, &total);
        if (idx < nbins) { bin_off[idx] = ex + carry; bin_cursor[idx] = ex + carry; }
        carry += total;
    }
    if (threadIdx.x == 0) bin_off[nbins] = carry;
}

// Tile of RS_PART_THREADS * PPT positions -> its entries, grouped by bin in shared memory, then written
// as one run per bin at a range claimed from the bin's cursor.  Dynamic shared memory:
// 2 * RS_MAX_BINS words + RS_PART_THREADS * PPT * W entries of 8 bytes.
template <int PPT>
__global__ void __launch_bounds__(RS_PART_THREADS)
rs_partition_kernel(MsmInput in, MsmGeom g, uint32_t nbins, uint32_t* bin_cursor, uint2* entries) {
    extern __shared__ uint32_t rs_sh[];
    uint32_t* cur = rs_sh;                       // counts, then tile-local cursors
    uint32_t* delta = rs_sh + RS_MAX_BINS;       // global position of local slot j of bin b = delta[b] + j
    uint2* stage = reinterpret_cast<uint2*>(rs_sh + 2 * RS_MAX_BINS);
    for (uint32_t b = threadIdx.x; b < RS_MAX_BINS; b += RS_PART_THREADS) cur[b] = 0;
    __syncthreads();
    const size_t tile0 = (size_t)blockIdx.x * (RS_PART_THREADS * PPT);
    uint32_t base_idx[PPT];
    bool alive[PPT];
    {
        DigitWalk dw[PPT];
#pragma unroll
        for (int p = 0; p < PPT; p++) {
            size_t i = tile0 + (size_t)p * RS_PART_THREADS + threadIdx.x;
            uint32_t s[8];
            base_idx[p] = 0;
            alive[p] = msm_prepare(in, g, i, nullptr, false, base_idx[p], s);
            dw[p].init(s, alive[p]);
        }
        for (uint32_t w = 0; w < g.W; w++) {
#pragma unroll
            for (int p = 0; p < PPT; p++) {
                bool neg;
                uint32_t key, tab;
                uint32_t d = dw[p].next(g, neg, key, tab);
                warp_agg_add_sh(cur, key >> RS_BIN_LOG, d != 0, false);
            }
        }
    }
    __syncthreads();
    // tile-local offsets of the bins (consecutive bins per thread) and one claim per non-empty bin
    uint32_t v[RS_MAX_BINS / RS_PART_THREADS], sum = 0;
#pragma unroll
    for (uint32_t k = 0; k < RS_MAX_BINS / RS_PART_THREADS; k++) {
        v[k] = cur[threadIdx.x * (RS_MAX_BINS / RS_PART_THREADS) + k];
        sum += v[k];
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan<(int)RS_PART_THREADS>(sum, &total);
#pragma unroll
    for (uint32_t k = 0; k < RS_MAX_BINS / RS_PART_THREADS; k++) {
        uint32_t b = threadIdx.x * (RS_MAX_BINS / RS_PART_THREADS) + k;
        if (v[k]) delta[b] = atomicAdd(bin_cursor + b, v[k]) - ex;
        cur[b] = ex;
        ex += v[k];
    }
    __syncthreads();
    {
        DigitWalk dw[PPT];
#pragma unroll
        for (int p = 0; p < PPT; p++) {              // the scalars again (L1 / L2 hits)
            size_t i = tile0 + (size_t)p * RS_PART_THREADS + threadIdx.x;
            uint32_t s[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            if (alive[p]) load_scalar(in.scalars + i * 8, s);
            dw[p].init(s, alive[p]);
        }
        for (uint32_t w = 0; w < g.W; w++) {
#pragma unroll
            for (int p = 0; p < PPT; p++) {
                bool neg;
                uint32_t key, tab;
                uint32_t d = dw[p].next(g, neg, key, tab);
                const bool live = d != 0;
                uint32_t pos = warp_agg_add_sh(cur, key >> RS_BIN_LOG, live, true);
                if (live) stage[pos] = make_uint2((base_idx[p] + tab) | (neg ? 0x80000000u : 0u), key);
            }
        }
    }
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < total; j += RS_PART_THREADS) {
        uint2 e = stage[j];
        entries[delta[e.y >> RS_BIN_LOG] + j] = e;
    }
}

// first bin whose range reaches past entry position s: largest b with bin_off[b] <= s
__device__ __forceinline__ uint32_t rs_find_bin(const uint32_t* bin_off, uint32_t nbins, uint32_t s) {
    uint32_t lo = 0, hi = nbins;                 // invariant: bin_off[lo] <= s < bin_off[hi]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(bin_off + mid) <= s) lo = mid; else hi = mid;
    }
    return lo;
}

// counts of one segment [s, e) of bin b's entries into sh[RS_BIN_BUCKETS] (zeroed here)
#define RS_UNROLL 4u
__device__ __forceinline__ void rs_segment_count(const uint2* entries, uint32_t s, uint32_t e, uint32_t* sh) {
    for (uint32_t i = threadIdx.x; i < RS_BIN_BUCKETS; i += RS_THREADS) sh[i] = 0;
    __syncthreads();
    for (uint32_t j0 = s; j0 < e; j0 += RS_UNROLL * RS_THREADS) {          // warp-uniform trip count
        uint32_t key[RS_UNROLL];
#pragma unroll
        for (uint32_t u = 0; u < RS_UNROLL; u++) {
            uint32_t j = j0 + u * RS_THREADS + threadIdx.x;
            key[u] = j < e ? __ldg(&entries[j].y) : 0xffffffffu;
        }
#pragma unroll
        for (uint32_t u = 0; u < RS_UNROLL; u++)
            warp_agg_add_sh(sh, key[u] & (RS_BIN_BUCKETS - 1u), key[u] != 0xffffffffu, false);
    }
    __syncthreads();
}

// Block k owns entries [k * chunk, (k + 1) * chunk) and walks the bins that range crosses.
__global__ void __launch_bounds__(RS_THREADS)
rs_bucket_hist_kernel(const uint2* entries, const uint32_t* bin_off, uint32_t nbins, uint32_t chunk,
                      uint32_t* hist) {
    __shared__ uint32_t sh[RS_BIN_BUCKETS];
    const uint32_t total = __ldg(bin_off + nbins);
    uint64_t s64 = (uint64_t)blockIdx.x * chunk;
    if (s64 >= total) return;
    uint32_t s = (uint32_t)s64;
    const uint32_t e = (uint32_t)(s64 + chunk < total ? s64 + chunk : total);
    uint32_t b = rs_find_bin(bin_off, nbins, s);
    while (s < e) {
        uint32_t be = __ldg(bin_off + b + 1);
        uint32_t se = be < e ? be : e;
        if (se > s) {
            rs_segment_count(entries, s, se, sh);
            for (uint32_t i = threadIdx.x; i < RS_BIN_BUCKETS; i += RS_THREADS)
                if (sh[i]) atomicAdd(hist + (size_t)b * RS_BIN_BUCKETS + i, sh[i]);
            __syncthreads();
            s = se;
        }
        if (s < e) b++;
    }
}

__global__ void __launch_bounds__(RS_THREADS)
rs_scatter_kernel(const uint2* entries, const uint32_t* bin_off, uint32_t nbins, uint32_t chunk,
                  uint32_t* cursor, uint32_t* sorted) {
    __shared__ uint32_t sh[RS_BIN_BUCKETS];
    const uint32_t total = __ldg(bin_off + nbins);
    uint64_t s64 = (uint64_t)blockIdx.x * chunk;
    if (s64 >= total) return;
    uint32_t s = (uint32_t)s64;
    const uint32_t e = (uint32_t)(s64 + chunk < total ? s64 + chunk : total);
    uint32_t b = rs_find_bin(bin_off, nbins, s);
    while (s < e) {
        uint32_t be = __ldg(bin_off + b + 1);
        uint32_t se = be < e ? be : e;
        if (se > s) {
            rs_segment_count(entries, s, se, sh);
            // one claim per non-empty bucket of the segment: sh[i] becomes this block's write cursor
            for (uint32_t i = threadIdx.x; i < RS_BIN_BUCKETS; i += RS_THREADS) {
                uint32_t c = sh[i];
                if (c) sh[i] = atomicAdd(cursor + (size_t)b * RS_BIN_BUCKETS + i, c);
            }
            __syncthreads();
            for (uint32_t j0 = s; j0 < se; j0 += RS_UNROLL * RS_THREADS) {
                uint2 en[RS_UNROLL];
#pragma unroll
                for (uint32_t u = 0; u < RS_UNROLL; u++) {
                    uint32_t j = j0 + u * RS_THREADS + threadIdx.x;
                    en[u] = j < se ? __ldg(entries + j) : make_uint2(0u, 0xffffffffu);
                }
#pragma unroll
                for (uint32_t u = 0; u < RS_UNROLL; u++) {
                    const bool live = en[u].y != 0xffffffffu;
                    uint32_t pos = warp_agg_add_sh(sh, en[u].y & (RS_BIN_BUCKETS - 1u), live, true);
                    if (live) sorted[pos] = en[u].x;
                }
            }
            __syncthreads();
            s = se;
        }
        if (s < e) b++;
    }
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
