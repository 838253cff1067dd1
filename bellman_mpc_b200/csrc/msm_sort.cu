// MSM stages 1-3 (host side): geometry, density prefix popcount, digit histogram, scans,
// counting-sort scatter.  Reference: src/multiexp.rs:191-223 (the per-window scan that this
// replaces) and :254-281 (window rule, density length).
#include <cmath>
#include <cstdlib>

#include "msm_sort_kernels.cuh"

namespace bmpc {

static uint32_t reference_window(size_t n) {  // multiexp.rs:267-271
    if (n < 32) return 3;
    return (uint32_t)std::ceil(std::log((double)(uint32_t)n));
}

static inline size_t scan_chunks_words(uint32_t n) {
    return (n + BMPC_SCAN_CHUNK - 1) / BMPC_SCAN_CHUNK + 2;
}

// Bits of the scalar that fall into the top window: 255 - c (W - 1).  When this is small every
// point drops its top digit into the same handful of buckets (c = 21: 3 bits -> 5 buckets with
// n / 5 points each; c dividing 255: only the carry, one bucket) -- hot atomics in the sort and
// thousands of partial sums to fold.  Window sizes with fewer than MIN_TOP_BITS are never chosen.
static inline uint32_t top_window_bits(uint32_t c) { return 255u - c * (255u / c); }
constexpr uint32_t MIN_TOP_BITS = 7;

// Cost model in point additions: n W bucket additions + 7 per bucket for the reduction (measured on
// B200: 0.30-0.36 ns per accumulated point, 2.3 ns per bucket: 2.5 ms for 2^19 buckets, 6.1 ms for 2^21).
static double window_cost(size_t n, uint32_t c, bool one_bucket_set) {
    uint32_t W = 255 / c + 1;
    double buckets = (double)(one_bucket_set ? 1u : W) * (double)(1u << (c - 1));
    return (double)n * W + 7.0 * buckets;
}
static uint32_t pick_window(size_t n, uint32_t lo, uint32_t hi, bool one_bucket_set) {
    uint32_t best = 0;
    double best_cost = 0;
    for (uint32_t c = lo; c <= hi; c++) {
        if (n >= (1u << 16) && top_window_bits(c) < MIN_TOP_BITS) continue;  // small n: harmless
        double cost = window_cost(n, c, one_bucket_set);
        if (!best || cost < best_cost) { best = c; best_cost = cost; }
    }
    return best ? best : lo;
}
static uint32_t plain_window(size_t n) {
    uint32_t lg = 0;
    while (((size_t)1 << (lg + 1)) <= n) lg++;
    uint32_t hi = lg > 8 ? lg - 3 : 5, lo = lg > 10 ? lg - 6 : 4;
    if (hi > 16) hi = 16;
    if (lo > hi) lo = hi;
    return pick_window(n, lo, hi, false);
}

// Window bits for precomputed tables: all windows share one bucket set, so the bucket count can
// grow until the (parallel) bucket reduction matters.
uint32_t msm_table_window(size_t n_bases) {
    uint32_t lg = 0;
    while (((size_t)1 << (lg + 1)) <= n_bases) lg++;
    uint32_t hi = lg > 5 ? lg - 1 : 4, lo = lg > 9 ? lg - 5 : 4;
    if (hi > 22) hi = 22;
    if (lo > hi) lo = hi;
    return pick_window(n_bases, lo, hi, true);
}

MsmPlan msm_make_plan(bmpc_ctx* ctx, const bmpc_bases* bases, size_t n_pos, bool has_density, size_t n_ref,
                      size_t n_dense) {
    // n_pos exponent positions, of which n (<= n_pos; the exact count when the caller knows it) are
    // dense and can contribute points: the geometry follows n, the density scan n_pos
    const size_t n = (has_density && n_dense && n_dense < n_pos) ? n_dense : n_pos;
    MsmPlan p;
    MsmGeom& g = p.g;
    uint32_t c;
    bool tables = bases->tab_W != 0;
    uint32_t ksets = 1;
    if (tables) {
        // The tables' window was sized for the whole vector; a multiexp over a few of its points
        // (create_proof's input multiexps: 16 exponents against a 2^22-point query vector) would pay
        // the reduction of 2^19 empty buckets (1.2 ms G1, 3.7 ms G2).  Such a call splits every table
        // window into k sub-windows of c' = c / k bits: sub-window s of table window t adds table t's
        // point into bucket set s (k sets of 2^(c'-1) buckets), and the k sets are folded with
        // (k-1) c' doublings -- instead of the 255 a table-less geometry needs.  k is the divisor of
        // the table window minimising additions + 7 per bucket + the doublings' latency (one cold
        // doubling on the serial tail ~ 22 us ~ 6 10^4 accumulated points).
        double best = window_cost(n, bases->tab_c, true);
        const bool subw = ctx->tune.subwindows != 0;
        for (uint32_t k = 2; k <= bases->tab_c / 2 && !ctx->tune_c && subw; k++) {
            if (bases->tab_c % k) continue;
            uint32_t cs = bases->tab_c / k;
            double cost = (double)n * (255 / cs + 1) + 7.0 * k * (double)(1u << (cs - 1)) + 6e4 * (k - 1) * cs;
            if (cost < best) { best = cost; ksets = k; }
        }
    }
    if (tables) c = bases->tab_c / ksets;
    else if (ctx->tune_c) c = (uint32_t)ctx->tune_c;
    else c = plain_window(n);
    if (c < 2) c = 2;
    if (c > 22) c = 22;
    g.c = c;
    g.W = 255 / c + 1;
    g.B = 1u << (c - 1);
    g.H = tables ? ksets : g.W;
    g.tab_stride = tables ? (uint32_t)bases->n : 0u;
    size_t avg = n * (g.W / g.H) / g.B;
    g.L = (uint32_t)(2 * avg < 64 ? 64 : 2 * avg);
    p.n = n_pos;
    p.has_density = has_density;
    // the reference picks its window from the length of the whole exponent vector; a shard of a
    // larger multiexp passes that length so that "which error wins" is decided as the reference does
    g.c_ref = reference_window(n_ref > n_pos ? n_ref : n_pos);
    g.top_skip = (254 / g.c_ref) * g.c_ref;
    p.nb = g.H * g.B;
    p.max_pairs = n * g.W;
    // Two-level partition sort (msm_sort_kernels.cuh: rs_*): pays from a few million pairs up, where the
    // one-pass scatter's random 4-byte stores stop being absorbed by L2.  BMPC_SORT_RADIX: 0 never,
    // 1 whenever the geometry allows it, -1 (default) from 2^22 pairs and 2^13 buckets up.
    {
        const int knob = ctx->tune.sort_radix;
        const bool fits = g.W <= RS_MAX_WINDOWS && (p.nb + RS_BIN_BUCKETS - 1) / RS_BIN_BUCKETS <= RS_MAX_BINS;
        p.radix_ok = fits && (knob > 0 || (knob < 0 && p.max_pairs >= ((size_t)1 << 22) && p.nb >= (1u << 13)));
    }
    // reduce: each thread owns S = 2^s_log consecutive buckets of one set (msm_reduce_kernel), blocks
    // of 256 threads.  S is the smallest power of two for which all H sets fit in ONE wave of
    // resident blocks (a second, nearly empty wave doubles the kernel time; 2 resident 256-thread
    // blocks per SM at <= 128 registers, 1 for G2) and a set has at most 256 blocks (the final
    // kernel folds a set's block results with one 256-thread block).  The kernel is latency-bound
    // (a dependent chain of 2 S + 17 cold additions per thread), so smaller S = shorter.
    {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
        size_t capacity = (size_t)sms * (bases->group == BMPC_G1 ? 2 : 1) * 256;
        uint32_t s_log = 0;
        while (s_log < g.c - 1 && (((size_t)g.H * g.B) >> s_log) > capacity) s_log++;
        for (int k = 0; k < ctx->tune.reduce_slog_add && s_log < g.c - 1; k++) s_log++;   // BMPC_REDUCE_SLOG_ADD: tuning knob
        // Threads per reduce block.  The block's suffix scan and tree cost log2(rblock) additions each
        // with every resident warp busy; the same levels cost a product's latency only in
        // msm_fold_kernel (single warps spread over the SMs), so the blocks are kept small and the
        // fold + final kernels take the upper levels.  BMPC_REDUCE_BLOCK: tuning knob.
        // Measured (2^19 buckets, combine + reduce + fold + final): G1 32 -> 2.38 ms, 64 -> 2.47, 128 -> 2.50,
        // 256 -> 2.58; G2 with the four-lane fold / final steps 32 -> 6.40 ms, 256 -> 7.43 (8.21 with
        // one-thread steps, where 256-thread blocks and no fold were the better choice).
        uint32_t rb = (bases->group == BMPC_G1 || ctx->tune.tail_quad != 0) ? 32 : 256;
        {
            uint32_t v = (uint32_t)ctx->tune.reduce_block;
            if (v == 32 || v == 64 || v == 128 || v == 256) rb = v;
        }
        while (s_log < g.c - 1 && (g.B >> s_log) > rb * BMPC_FOLD_GROUP * 256u) s_log++;
        p.s_log = s_log;
        p.tpw = g.B >> s_log;
        p.rblock = p.tpw < rb ? p.tpw : rb;
        p.nblk = p.tpw / p.rblock;
    }
    msm_plan_sizes(p);
    return p;
}

void msm_plan_sizes(MsmPlan& p) {
    // pair mode pads every bucket to an even number of entries (one pad entry per odd bucket)
    const size_t entries = p.max_pairs + (p.pairs ? p.nb : 0);
    p.max_tasks = entries / p.g.L + p.nb + 1;
    size_t b = 0;
    size_t nw32 = (p.n + 31) / 32;
    if (p.has_density) b += ws_need(nw32 + 1, 4) * 2 + ws_need(scan_chunks_words((uint32_t)nw32), 4);
    b += ws_need(p.nb + 1, 4) * 5;  // hist, off, toff, cursor, heavy
    b += ws_need(scan_chunks_words(p.nb), 4);
    b += ws_need(entries + 2, 4);   // sorted
    b += ws_need(64, 4) + ws_need(256, 4);
    b += ws_need(p.max_tasks, 16);     // task descriptors
    p.radix = p.radix_ok && !p.pairs;
    if (p.radix) b += ws_need(p.max_pairs + 2, 8) + 3 * ws_need(RS_MAX_BINS + 1, 4);   // entries, bin hist / off / cursor
    p.sort_bytes = b + 4096;
}

// out[n+1] = exclusive scan of xform(in); `chunks` scratch of scan_chunks_words(n) words
static int run_scan(bmpc_ctx* ctx, const uint32_t* in, uint32_t n, uint32_t L, uint32_t* chunks,
                    uint32_t* out, cudaStream_t st) {
    uint32_t nchunks = (n + BMPC_SCAN_CHUNK - 1) / BMPC_SCAN_CHUNK;
    if (nchunks == 0) nchunks = 1;
    LAUNCH(ctx, scan_phase1_kernel, nchunks, BMPC_SCAN_THREADS, 0, st, in, n, L, chunks);
    LAUNCH(ctx, scan_phase2_kernel, 1, BMPC_SCAN_THREADS, 0, st, chunks, nchunks);
    LAUNCH(ctx, scan_phase3_kernel, nchunks, BMPC_SCAN_THREADS, 0, st, in, n, L, chunks, nchunks, out);
    return BMPC_OK;
}

int msm_sort_run(bmpc_ctx* ctx, const MsmPlan& p, const bmpc_bases* bases, size_t base_offset,
                 const uint32_t* d_scalars, size_t n, const uint32_t* d_density, uint32_t* d_flags,
                 MsmSorted* out, cudaStream_t st) {
    const MsmGeom& g = p.g;
    ProfScope ps(ctx, BMPC_PROF_MSM_SORT, st);
    MsmInput in;
    in.scalars = d_scalars;
    in.n = n;
    in.density = d_density;
    in.word_prefix = nullptr;
    in.base_offset = (uint32_t)base_offset;
    in.bases_len = (uint32_t)bases->n;
    in.inf_bitmap = bases->d_inf;
    CK(cudaMemsetAsync(d_flags, 0, 16, st));
    if (d_density) {
        uint32_t nw32 = (uint32_t)((n + 31) / 32);
        uint32_t* pc = ws_take<uint32_t>(ctx, nw32 + 1);
        uint32_t* prefix = ws_take<uint32_t>(ctx, nw32 + 1);
        uint32_t* chunks = ws_take<uint32_t>(ctx, scan_chunks_words(nw32));
        if (!pc || !prefix || !chunks) {
            ctx->err = "msm workspace carve failed (density)";
            return BMPC_ERR_INVALID;
        }
        LAUNCH(ctx, popc_words_kernel, (nw32 + 255) / 256, 256, 0, st, d_density, nw32, (uint32_t)n, pc);
        int rc = run_scan(ctx, pc, nw32, 0, chunks, prefix, st);
        if (rc) return rc;
        in.word_prefix = prefix;
    }
    uint32_t* hist = ws_take<uint32_t>(ctx, p.nb + 1);
    uint32_t* off = ws_take<uint32_t>(ctx, p.nb + 1);
    uint32_t* toff = ws_take<uint32_t>(ctx, p.nb + 1);
    uint32_t* cursor = ws_take<uint32_t>(ctx, p.nb + 1);
    uint32_t* heavy = ws_take<uint32_t>(ctx, p.nb + 1);
    uint32_t* chunks = ws_take<uint32_t>(ctx, scan_chunks_words(p.nb));
    const size_t sorted_entries = p.max_pairs + (p.pairs ? p.nb : 0) + 2;
    uint32_t* sorted = ws_take<uint32_t>(ctx, sorted_entries);
    uint32_t* heavy_count = ws_take<uint32_t>(ctx, 64);
    uint32_t* bins = ws_take<uint32_t>(ctx, 256);
    uint4* desc = ws_take<uint4>(ctx, p.max_tasks);
    if (!bins || !desc) {
        ctx->err = "msm workspace carve failed (tasks)";
        return BMPC_ERR_INVALID;
    }
    if (!hist || !off || !toff || !cursor || !heavy || !chunks || !sorted || !heavy_count) {
        ctx->err = "msm workspace carve failed";
        return BMPC_ERR_INVALID;
    }
    CK(cudaMemsetAsync(hist, 0, (size_t)(p.nb + 1) * 4, st));
    CK(cudaMemsetAsync(heavy_count, 0, 4, st));
    uint32_t nblocks = (uint32_t)((n + 255) / 256);
    // ---- two-level partition sort: bins of RS_BIN_BUCKETS buckets, then the buckets of a bin
    uint2* entries = nullptr;
    uint32_t *bin_hist = nullptr, *bin_off = nullptr, *bin_cursor = nullptr;
    const uint32_t nbins = (p.nb + RS_BIN_BUCKETS - 1) / RS_BIN_BUCKETS;
    uint32_t rs_chunk = 0, rs_blocks = 0;
    if (p.radix) {
        entries = ws_take<uint2>(ctx, p.max_pairs + 2);
        bin_hist = ws_take<uint32_t>(ctx, RS_MAX_BINS + 1);
        bin_off = ws_take<uint32_t>(ctx, RS_MAX_BINS + 1);
        bin_cursor = ws_take<uint32_t>(ctx, RS_MAX_BINS + 1);
        if (!entries || !bin_hist || !bin_off || !bin_cursor) {
            ctx->err = "msm workspace carve failed (partition sort)";
            return BMPC_ERR_INVALID;
        }
        CK(cudaMemsetAsync(bin_hist, 0, (size_t)(RS_MAX_BINS + 1) * 4, st));
        // tile of 1024 (or, for many windows, 512) positions: count kernel 256 threads x 4 (2), partition
        // kernel 512 threads x 2 (1)
        const int ppt = g.W <= RS_PPT4_MAX_WINDOWS ? 4 : 2;
        const uint32_t tile = RS_THREADS * ppt;
        const uint32_t tiles = (uint32_t)((n + tile - 1) / tile);
        const size_t part_smem = (size_t)2 * RS_MAX_BINS * 4 + (size_t)tile * g.W * 8;
        if (ppt == 4) {
            CK(cudaFuncSetAttribute(rs_partition_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)part_smem));
            LAUNCH(ctx, rs_bin_count_kernel<4>, tiles, RS_THREADS, nbins * 4, st, in, g, nbins, bin_hist, d_flags);
            LAUNCH(ctx, rs_bin_scan_kernel, 1, BMPC_SCAN_THREADS, 0, st, bin_hist, nbins, bin_off, bin_cursor);
            LAUNCH(ctx, rs_partition_kernel<2>, tiles, RS_PART_THREADS, part_smem, st, in, g, nbins, bin_cursor, entries);
        } else {
            CK(cudaFuncSetAttribute(rs_partition_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)part_smem));
            LAUNCH(ctx, rs_bin_count_kernel<2>, tiles, RS_THREADS, nbins * 4, st, in, g, nbins, bin_hist, d_flags);
            LAUNCH(ctx, rs_bin_scan_kernel, 1, BMPC_SCAN_THREADS, 0, st, bin_hist, nbins, bin_off, bin_cursor);
            LAUNCH(ctx, rs_partition_kernel<1>, tiles, RS_PART_THREADS, part_smem, st, in, g, nbins, bin_cursor, entries);
        }
        // Entries per block of the bucket-side kernels.  A bucket's slice of `sorted` is completed by all
        // the chunks of its bin, so its sectors must stay in L2 for as long as a chunk takes: what is in
        // flight -- resident blocks x chunk x 4 bytes of output -- has to stay well below L2.  Measured at
        // 2^24 points (sort stage in ms): 2^16 7.5, 2^14 3.7, 2^13 3.3, 2^12 3.0, 2^11 no better; with
        // 2^10 buckets per bin a chunk of 2^12 still brings four entries per bucket, so one claim serves
        // a run.  BMPC_RS_CHUNK_LOG: tuning knob.
        rs_chunk = 1u << 12;
        if (ctx->tune.rs_chunk_log >= 10 && ctx->tune.rs_chunk_log <= 20) rs_chunk = 1u << ctx->tune.rs_chunk_log;
        rs_blocks = (uint32_t)((p.max_pairs + rs_chunk - 1) / rs_chunk);
        if (!rs_blocks) rs_blocks = 1;
        LAUNCH(ctx, rs_bucket_hist_kernel, rs_blocks, RS_THREADS, 0, st, (const uint2*)entries, (const uint32_t*)bin_off,
               nbins, rs_chunk, hist);
    } else {
        LAUNCH(ctx, msm_count_kernel, nblocks, 256, 0, st, in, g, hist, d_flags);
    }
    // pair mode: bucket offsets with every count rounded up to even, the pad slots keep BMPC_PAIR_PAD
    int rc = run_scan(ctx, hist, p.nb, p.pairs ? BMPC_SCAN_EVEN : 0, chunks, off, st);
    if (rc) return rc;
    if (p.pairs) CK(cudaMemsetAsync(sorted, 0xff, sorted_entries * 4, st));
    rc = run_scan(ctx, hist, p.nb, g.L, chunks, toff, st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(cursor, off, (size_t)p.nb * 4, cudaMemcpyDeviceToDevice, st));
    if (p.radix)
        LAUNCH(ctx, rs_scatter_kernel, rs_blocks, RS_THREADS, 0, st, (const uint2*)entries, (const uint32_t*)bin_off,
               nbins, rs_chunk, cursor, sorted);
    else
        LAUNCH(ctx, msm_scatter_kernel, nblocks, 256, 0, st, in, g, cursor, sorted);
    LAUNCH(ctx, msm_find_heavy_kernel, (p.nb + 255) / 256, 256, 0, st, toff, p.nb, heavy, heavy_count);
    CK(cudaMemsetAsync(bins, 0, 256 * 4, st));
    LAUNCH(ctx, task_bin_count_kernel, (p.nb + 255) / 256, 256, 0, st, off, toff, p.nb, g.L, bins);
    LAUNCH(ctx, task_bin_scan_kernel, 1, 1, 0, st, bins);
    LAUNCH(ctx, task_desc_kernel, (p.nb + 255) / 256, 256, 0, st, off, toff, p.nb, g.L, bins, desc);
    out->desc = desc;
    out->ntasks = toff + p.nb;
    out->sorted = sorted;
    out->off = off;
    out->toff = toff;
    out->heavy = heavy;
    out->heavy_count = heavy_count;
    out->nsorted = off + p.nb;
    return BMPC_OK;
}

int msm_pairs_prepare(bmpc_ctx* ctx, const MsmPlan& p, const MsmSorted& s, uint32_t* csums, uint32_t* totals,
                      uint32_t* pairoff, uint2* lists, uint32_t* fin, cudaStream_t st) {
    const uint32_t R = p.pair_R;
    static_assert(BMPC_SCAN_CHUNK == 1024, "pair_cstride assumes scan chunks of 1024 tasks");
    const uint32_t nchunks = (uint32_t)((p.max_tasks + BMPC_SCAN_CHUNK - 1) / BMPC_SCAN_CHUNK);
    dim3 sgrid(nchunks, R - 1);
    LAUNCH(ctx, pair_scan_phase1_kernel, sgrid, BMPC_SCAN_THREADS, 0, st, s.desc, s.ntasks, csums, p.pair_cstride);
    LAUNCH(ctx, pair_scan_phase2_kernel, R - 1, BMPC_SCAN_THREADS, 0, st, csums, p.pair_cstride, nchunks, totals);
    LAUNCH(ctx, pair_scan_phase3_kernel, sgrid, BMPC_SCAN_THREADS, 0, st, s.desc, s.ntasks, (const uint32_t*)csums,
           p.pair_cstride, pairoff, p.pair_stride);
    PairLayout lay;
    lay.R = R;
    for (uint32_t r = 0; r <= BMPC_PAIR_MAX_ROUNDS; r++) {
        lay.out_base[r] = r <= R ? p.pair_out[r] : 0;
        lay.list_off[r] = r <= R ? p.pair_list[r] : 0;
    }
    LAUNCH(ctx, pair_build_kernel, (uint32_t)((p.max_tasks + 127) / 128), 128, 0, st, s.desc, s.ntasks, lay,
           (const uint32_t*)pairoff, p.pair_stride, lists, fin);
    return BMPC_OK;
}

}  // namespace bmpc
