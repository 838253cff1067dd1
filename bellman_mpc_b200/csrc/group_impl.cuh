// Host-side launchers for the curve-typed kernels, instantiated once per group
// (group_g1.cu: F = Fp, group_g2.cu: F = Fp2).
#pragma once
#include <cstdlib>

#include "internal.h"
#include "group_kernels.cuh"
#include "msm_affine.cuh"
#include "msm_pairs.cuh"

namespace bmpc {

template <class F>
using AffKernel = void (*)(const Affine<F>*, const uint32_t*, const uint4*, const uint32_t*, XYZZ<F>*, Affine<F>*,
                           uint32_t, uint32_t, uint32_t, uint32_t);
template <class F>
static AffKernel<F> aff_kernel(uint32_t K, uint32_t minb, uint32_t blk = 128) {
    // 256-thread blocks (BMPC_AFF_BLOCKDIM=256, experiment): half as many block inversions per SM,
    // two resident blocks instead of four
    if constexpr (sizeof(F) == sizeof(Fp)) {
        if (blk == 512) return msm_accumulate_affine_kernel<F, 384, 1, 512>;
        if (blk == 256) return msm_accumulate_affine_kernel<F, 384, 2, 256>;
    } else {
        if (blk == 256) return msm_accumulate_affine_kernel<F, 384, 1, 256>;   // G2: one block per SM at 255 registers
    }
    if (K == 384) {
        if (minb == 4) return msm_accumulate_affine_kernel<F, 384, 4>;
        if (minb == 3) return msm_accumulate_affine_kernel<F, 384, 3>;
        return msm_accumulate_affine_kernel<F, 384, 1>;
    }
    return minb == 4 ? msm_accumulate_affine_kernel<F, 128, 4> : msm_accumulate_affine_kernel<F, 128, 1>;
}

// ---- round-based pair accumulation (msm_pairs.cuh) -------------------------------------------
template <class F>
struct PairCfg;
template <>
struct PairCfg<Fp> {       // 61 KB of staging per block: three 128-thread blocks per SM at <= 168 registers
    static constexpr int BLK = 128, MINB = 3;
    static constexpr bool PK = true;
};
template <>
struct PairCfg<Fp2> {      // 96 KB of staging per block (prefix product read directly): two blocks per SM
    static constexpr int BLK = 128, MINB = 2;
    static constexpr bool PK = false;
};
template <class F>
static size_t pair_smem_bytes() {
    typedef PairStage<F, PairCfg<F>::BLK, PairCfg<F>::PK> Stage;
    size_t tree = 4 * (size_t)PairCfg<F>::BLK * sizeof(F);
    return Stage::BYTES > tree ? Stage::BYTES : tree;
}
template <class F, bool R0>
static void (*pair_kernel())(PairArgs<F>) {
    return msm_pair_round_kernel<F, R0, PairCfg<F>::BLK, PairCfg<F>::MINB, PairCfg<F>::PK>;
}

// Decides whether this multiexp accumulates its buckets in pair rounds and lays the rounds out.
// BMPC_ACC_PAIRS = 0 never, 1 always (tests run every accumulate kernel on the same inputs); default:
// from BMPC_PAIR_MIN_ENTRIES (position, window) entries up, where the rounds fill the GPU.
template <class F>
static bool plan_pairs(bmpc_ctx* ctx, MsmPlan& p) {
    const int mode = ctx->tune.acc_pairs;
    const size_t min_entries = ctx->tune.pair_min_entries;
    const uint32_t kmax_env = ctx->tune.pair_k > 0 ? (uint32_t)ctx->tune.pair_k : 0;
    if (mode == 0) return false;
    if (mode != 1 && p.max_pairs < min_entries) return false;
    // slices of at most L = 2^R entries, L >= twice the mean bucket (most buckets are one slice)
    size_t avg = p.max_pairs / (p.nb ? p.nb : 1);
    uint32_t R = 6;
    while (R < BMPC_PAIR_MAX_ROUNDS && ((size_t)1 << R) < 2 * avg) R++;
    MsmPlan q = p;
    q.pairs = true;
    q.g.L = 1u << R;
    q.pair_R = R;
    msm_plan_sizes(q);
    const size_t entries = q.max_pairs + q.nb;               // padded entry bound
    size_t ob = 0, lo = 0;
    for (uint32_t r = 0; r <= R; r++) {
        q.pair_out[r] = (uint32_t)ob;
        q.pair_list[r] = (uint32_t)lo;
        size_t pmax = r == 0 ? entries / 2 + 1 : (entries >> (r + 1)) + q.max_tasks + 1;
        ob += pmax;
        if (r >= 1) lo += pmax;
        if (ob >= ((size_t)1 << 32) - 1) return false;       // pool indices are 32-bit
    }
    int sms = 148, occ = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    const size_t smem = pair_smem_bytes<F>();
    cudaFuncSetAttribute(pair_kernel<F, true>(), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(pair_kernel<F, false>(), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pair_kernel<F, true>(), PairCfg<F>::BLK, smem);
    if (occ < 1) return false;
    q.pair_block = PairCfg<F>::BLK;
    q.pair_minb = (uint32_t)occ;
    q.pair_blocks = (uint32_t)(sms * occ);
    q.pair_kmax = kmax_env ? kmax_env : 256;
    q.pair_stride = (uint32_t)align_up(q.max_tasks, 64);
    q.pair_cstride = (uint32_t)align_up((q.max_tasks + 1023) / 1024 + 2, 64);   // scan chunks of 1024 tasks
    q.affine = false;
    p = q;
    return true;
}
template <class F>
static size_t pair_bytes(const MsmPlan& p) {
    const uint32_t R = p.pair_R;
    return ws_need(p.pair_out[R], sizeof(Affine<F>)) + ws_need((size_t)p.pair_list[R] + 1, 8) +
           ws_need((size_t)(R ? R - 1 : 0) * p.pair_stride + 1, 4) + ws_need((size_t)(R ? R - 1 : 0) * p.pair_cstride + 1, 4) +
           ws_need(64, 4) + ws_need(p.max_tasks, 4) +
           ws_need((size_t)p.pair_blocks * p.pair_block * p.pair_kmax, sizeof(F));
}

// Batched-affine accumulation pays when there are enough bucket slices to give every resident
// thread a job of >= 2 slices (fewer: the inversion is not amortised and the XYZZ kernel wins).
// BMPC_ACC_AFFINE=0 disables it, =1 forces it (tests run both paths on the same inputs).
template <class F>
void GroupOps<F>::plan_affine(bmpc_ctx* ctx, MsmPlan& p) {
    const bmpc_tuning& tn = ctx->tune;
    int mode = tn.acc_affine;
    p.affine = false;
    if (plan_pairs<F>(ctx, p)) return;
    if (mode == 0) return;
    int sms = 148, occ = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    // tuning knobs: threads per block (power of two <= 128), additions per inversion (128 | 384)
    // G1: 256-thread blocks (two per SM) halve the block inversions of four 128-thread blocks: 2^24
    // 59.78 -> 58.36 ms, 2^21 9.04 -> 8.97 ms, create_proof 2^22 0.1263 -> 0.1231 s; one 512-thread
    // block per SM pays only for the largest bucket sets (2^24: 57.86 ms, but 2^21: 9.43 ms).
    // G2 stays at 128 (168 registers, 3 blocks per SM).
    uint32_t blk = sizeof(F) == sizeof(Fp) ? (p.nb >= (1u << 21) ? 512 : 256) : 128;
    if (tn.aff_blockdim) {
        uint32_t v = (uint32_t)tn.aff_blockdim;
        if (v == 32 || v == 64 || v == 128 || v == 256 || (v == 512 && sizeof(F) == sizeof(Fp))) blk = v;
    }
    // measured at 2^24 (G1): K = 384 / 128 registers (4 blocks per SM) 59.5 ms, K = 128 61.5 ms,
    // 172 registers (2 blocks) 80 ms; XYZZ kernel 73.2 ms.  G2 at 2^22: 168 registers (3 blocks per
    // SM, 0.9 KB of spills) 60.3 ms, 252 registers (2 blocks) 65.0 ms, XYZZ kernel 72.0 ms.
    p.aff_K = tn.aff_ksel == 128 ? 128 : 384;
    // G2, round 3: with the Fp products of every Fp2 product as calls (group_g2.cu: BMPC_FP2_CALLS) the
    // kernel body is a fifth of the inlined 197 KB and the instruction cache stops being the limit; two
    // blocks per SM at 255 registers (no spills) then beat three at 168 (2^21 points: inlined 29.4 ms,
    // calls at 168 registers 27.6, calls at 255 registers 25.0, calls at 128 registers 31.1).
    p.aff_minb = sizeof(F) == sizeof(Fp) ? 4 : 1;
    if (tn.aff_minb) {
        int v = tn.aff_minb;
        p.aff_minb = (v == 4 || v == 3) ? (uint32_t)v : 1u;
    }
    p.aff_block = blk;
    size_t smem = 4 * (size_t)blk * sizeof(F);
    auto kern = aff_kernel<F>(p.aff_K, p.aff_minb, blk);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, (int)blk, smem);
    if (occ < 1) occ = 1;
    size_t resident = (size_t)sms * occ * blk;
    size_t gmax = BMPC_AFF_G;                      // BMPC_AFF_GMAX: tuning knob (slices per job)
    if (tn.aff_gmax >= 1 && (size_t)tn.aff_gmax < gmax) gmax = (size_t)tn.aff_gmax;
    // Slices per job.  Jobs are handed out in waves of `resident` threads; with a single wave the
    // kernel's second half runs on a half-empty GPU (threads finish at different times and nothing
    // refills them), so G is chosen among the values that give at least TWO waves, maximising the
    // fill of the last one: nb / (ceil(nb / (G resident)) G resident), nb ~ number of slices.
    // Measured at 2^22 (G1, 2^19 buckets): G = 4 (2 waves) 18.1 ms, G = 7 (1 wave) 20.9 ms, XYZZ 20.1 ms.
    size_t G = 0;
    double best = 0;
    // BMPC_AFF_WHOLE_WAVES (default on): the kernel deals the slices over a whole number of waves
    // (njobs = waves x resident threads), so the fill of the last wave no longer depends on G and G
    // only sets the amortisation: the largest G <= gmax that still leaves two waves.
    p.aff_whole_waves = tn.aff_whole_waves != 0;
    if (p.aff_whole_waves) {
        size_t per_thread = (p.nb + resident - 1) / resident;      // slices per resident thread
        size_t waves = tn.aff_waves >= 1 ? (size_t)tn.aff_waves : 2;
        G = (per_thread + waves - 1) / waves;                       // `waves` (2) waves of jobs this size
        if (G > gmax) G = gmax;
        if (G < 2) G = 2;
        if (p.nb <= 2 * resident) G = 0;                            // no two waves of jobs: XYZZ kernel
    } else {
        for (size_t cand = 2; cand <= gmax; cand++) {
            size_t waves = (p.nb + cand * resident - 1) / (cand * resident);
            double eff = (double)p.nb / ((double)waves * cand * resident);
            if (waves >= 2 && eff >= best - 1e-9) { best = eff; G = cand; }
        }
    }
    // Automatic choice (BMPC_ACC_AFFINE unset): where it was measured to win -- enough slices for two
    // waves of jobs of >= 2 slices (G1: 2^19 buckets and up, i.e. 2^21-point multiexps with window
    // tables: 18.1 vs 20.1 ms at 2^22, 31.3 vs 36.8 at 2^23, 59.5 vs 73.2 at 2^24; G2 69.7 vs 74.4 at 2^22).
    if (G < 2) {
        if (mode != 1) return;
        G = tn.aff_force_g ? (size_t)tn.aff_force_g : 2;
        if (G < 1) G = 1;
        if (G > BMPC_AFF_G) G = BMPC_AFF_G;
    }
    size_t jobs = (p.max_tasks + G - 1) / G;
    size_t blocks = (jobs + blk - 1) / blk;
    if (blocks > (size_t)sms * occ) blocks = (size_t)sms * occ;
    p.affine = true;
    p.aff_G = (uint32_t)G;
    p.aff_blocks = (uint32_t)blocks;
    p.aff_HA = (p.g.L + 1) / 2;
    p.aff_HB = (p.aff_HA + 1) / 2;
}

template <class F>
size_t GroupOps<F>::curve_bytes(const MsmPlan& p) {
    if (p.pairs) return curve_bytes_xyzz(p) + pair_bytes<F>(p);
    if (p.affine)
        return curve_bytes_xyzz(p) + ws_need((size_t)p.aff_blocks * p.aff_block * p.aff_G * (p.aff_HA + p.aff_HB), sizeof(Affine<F>));
    return curve_bytes_xyzz(p);
}
template <class F>
size_t GroupOps<F>::curve_bytes_xyzz(const MsmPlan& p) {
    return ws_need(p.max_tasks, sizeof(XYZZ<F>)) + 2 * ws_need((size_t)p.g.H * p.nblk, sizeof(XYZZ<F>)) +
           2 * ws_need((size_t)p.g.H * (p.nblk / BMPC_FOLD_GROUP + 1), sizeof(XYZZ<F>)) +
           2 * ws_need((size_t)p.g.H * (p.nblk / (BMPC_FOLD_GROUP * BMPC_FOLD_GROUP) + 2), sizeof(XYZZ<F>)) +
           ws_need(p.g.H, sizeof(XYZZ<F>)) + ws_need(64, 4) + 1024;
}

template <class F>
int GroupOps<F>::msm_finish(bmpc_ctx* ctx, const MsmPlan& p, const bmpc_bases* bases, const MsmSorted& s,
                            int mode, uint8_t* d_out_bytes, void* d_out_xyzz, cudaStream_t st) {
    const MsmGeom& g = p.g;
    XYZZ<F>* partials = ws_take<XYZZ<F>>(ctx, p.max_tasks);
    XYZZ<F>* blk_V = ws_take<XYZZ<F>>(ctx, (size_t)g.H * p.nblk);
    XYZZ<F>* blk_R = ws_take<XYZZ<F>>(ctx, (size_t)g.H * p.nblk);
    // block results are folded in groups of BMPC_FOLD_GROUP before the final kernel when a set has
    // more of them than the final kernel's one block takes (msm_fold_kernel)
    const uint32_t groups = p.nblk > BMPC_FINAL_THREADS ? p.nblk / BMPC_FOLD_GROUP : 0;
    XYZZ<F>* grp_V = ws_take<XYZZ<F>>(ctx, (size_t)g.H * (p.nblk / BMPC_FOLD_GROUP + 1));
    XYZZ<F>* grp_R = ws_take<XYZZ<F>>(ctx, (size_t)g.H * (p.nblk / BMPC_FOLD_GROUP + 1));
    XYZZ<F>* grp2_V = ws_take<XYZZ<F>>(ctx, (size_t)g.H * (p.nblk / (BMPC_FOLD_GROUP * BMPC_FOLD_GROUP) + 2));
    XYZZ<F>* grp2_R = ws_take<XYZZ<F>>(ctx, (size_t)g.H * (p.nblk / (BMPC_FOLD_GROUP * BMPC_FOLD_GROUP) + 2));
    XYZZ<F>* win = ws_take<XYZZ<F>>(ctx, g.H);
    uint32_t* ticket = ws_take<uint32_t>(ctx, 64);
    if (!partials || !blk_V || !blk_R || !grp_V || !grp_R || !grp2_V || !grp2_R || !win || !ticket) {
        ctx->err = "msm workspace carve failed (curve)";
        return BMPC_ERR_INVALID;
    }
    const Affine<F>* pts = reinterpret_cast<const Affine<F>*>(bases->d_points);
    uint32_t ablocks = (uint32_t)((p.max_tasks + 127) / 128);
    if (p.pairs) {
        const uint32_t R = p.pair_R;
        Affine<F>* pool = ws_take<Affine<F>>(ctx, p.pair_out[R]);
        uint2* lists = ws_take<uint2>(ctx, (size_t)p.pair_list[R] + 1);
        uint32_t* pairoff = ws_take<uint32_t>(ctx, (size_t)(R - 1) * p.pair_stride + 1);
        uint32_t* csums = ws_take<uint32_t>(ctx, (size_t)(R - 1) * p.pair_cstride + 1);
        uint32_t* totals = ws_take<uint32_t>(ctx, 64);
        uint32_t* fin = ws_take<uint32_t>(ctx, p.max_tasks);
        F* pre = ws_take<F>(ctx, (size_t)p.pair_blocks * p.pair_block * p.pair_kmax);
        if (!pool || !lists || !pairoff || !csums || !totals || !fin || !pre) {
            ctx->err = "msm workspace carve failed (pair rounds)";
            return BMPC_ERR_INVALID;
        }
        ProfScope ps(ctx, BMPC_PROF_MSM_ACCUMULATE, st);
        // per-round task scans -> dense pair lists of rounds 1 .. R-1 (msm_sort.cu)
        int rcp = msm_pairs_prepare(ctx, p, s, csums, totals, pairoff, lists, fin, st);
        if (rcp) return rcp;
        const size_t smem = pair_smem_bytes<F>();
        for (uint32_t r = 0; r < R; r++) {
            PairArgs<F> a;
            a.tables = pts;
            a.out = pool;
            a.list = r == 0 ? reinterpret_cast<const uint2*>(s.sorted) : lists + p.pair_list[r];
            a.count_p = r == 0 ? s.nsorted : totals + r;
            a.count_shift = r == 0 ? 1u : 0u;
            a.out_base = p.pair_out[r];
            a.pre = pre;
            a.kmax = p.pair_kmax;
            if (r == 0) LAUNCH(ctx, (pair_kernel<F, true>()), p.pair_blocks, p.pair_block, smem, st, a);
            else LAUNCH(ctx, (pair_kernel<F, false>()), p.pair_blocks, p.pair_block, smem, st, a);
        }
        LAUNCH(ctx, msm_pair_collect_kernel<F>, ablocks, 128, 0, st, (const Affine<F>*)pool, (const uint32_t*)fin,
               s.desc, s.ntasks, partials);
    } else if (p.affine) {
        Affine<F>* scratch = ws_take<Affine<F>>(ctx, (size_t)p.aff_blocks * p.aff_block * p.aff_G * (p.aff_HA + p.aff_HB));
        if (!scratch) {
            ctx->err = "msm workspace carve failed (affine scratch)";
            return BMPC_ERR_INVALID;
        }
        ProfScope ps(ctx, BMPC_PROF_MSM_ACCUMULATE, st);
        size_t smem = 4 * (size_t)p.aff_block * sizeof(F);
        auto kern = aff_kernel<F>(p.aff_K, p.aff_minb, p.aff_block);
        LAUNCH(ctx, kern, p.aff_blocks, p.aff_block, smem, st, pts, s.sorted, s.desc, s.ntasks, partials, scratch,
               p.aff_HA, p.aff_HB, p.aff_G, (uint32_t)p.aff_whole_waves);
    } else {
        ProfScope ps(ctx, BMPC_PROF_MSM_ACCUMULATE, st);
        // BMPC_ACC_COMPACT=1 selects the variant whose field products are calls (smaller code)
        const bool compact = ctx->tune.acc_compact != 0;
        if (compact)
            LAUNCH(ctx, (msm_accumulate_kernel<F, true>), ablocks, 128, 0, st, pts, s.sorted, s.desc, s.ntasks, partials);
        else
            LAUNCH(ctx, (msm_accumulate_kernel<F, false>), ablocks, 128, 0, st, pts, s.sorted, s.desc, s.ntasks, partials);
    }
    ProfScope ps_tail(ctx, BMPC_PROF_MSM_REDUCE, st);
    {
        size_t smem = 128 * sizeof(XYZZ<F>);
        CK(cudaFuncSetAttribute(msm_combine_heavy_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LAUNCH(ctx, msm_combine_heavy_kernel<F>, 148 * 3, 128, smem, st, s.toff, s.heavy, s.heavy_count, partials);
    }
    {
        size_t smem = (size_t)p.rblock * sizeof(XYZZ<F>);
        CK(cudaFuncSetAttribute(msm_reduce_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(256 * sizeof(XYZZ<F>))));
        dim3 grid(p.nblk, g.H);
        LAUNCH(ctx, msm_reduce_kernel<F>, grid, p.rblock, smem, st, (const XYZZ<F>*)partials, s.toff, g.B, p.s_log,
               blk_V, blk_R);
    }
    {
        CK(cudaMemsetAsync(ticket, 0, 4, st));
        uint32_t m_log = p.s_log, rb = p.rblock;
        while (rb > 1) { m_log++; rb >>= 1; }
        const XYZZ<F>*fin_V = blk_V, *fin_R = blk_R;
        uint32_t fin_n = p.nblk;
        // four lanes per element (quad.cuh): fold in groups of 32 until at most 64 results are left, then
        // the final step (2^19 buckets: G1 2.36 -> 2.12 ms, G2 8.21 -> 6.40 ms for combine + reduce + fold +
        // final); BMPC_TAIL_QUAD=0: one thread per element, one fold when a set has more than 256 results
        const bool quad = ctx->tune.tail_quad != 0 && (p.nblk & (p.nblk - 1)) == 0;
        if (quad) {
            XYZZ<F>* outV[2] = {grp_V, grp2_V};
            XYZZ<F>* outR[2] = {grp_R, grp2_R};
            int lvl = 0;
            while (fin_n > 64) {
                const uint32_t ng = fin_n / BMPC_FOLD_GROUP;
                dim3 fgrid(ng, g.H);
                LAUNCH(ctx, msm_fold_quad_kernel<F>, fgrid, 4 * BMPC_FOLD_GROUP, BMPC_FOLD_GROUP * sizeof(XYZZ<F>), st,
                       fin_V, fin_R, fin_n, m_log, outV[lvl & 1], outR[lvl & 1]);
                for (uint32_t q = BMPC_FOLD_GROUP; q > 1; q >>= 1) m_log++;
                fin_V = outV[lvl & 1]; fin_R = outR[lvl & 1]; fin_n = ng;
                lvl++;
            }
            uint32_t cnt = 8;
            while (cnt < fin_n) cnt <<= 1;
            LAUNCH(ctx, msm_final_quad_kernel<F>, g.H, 4 * cnt, cnt * sizeof(XYZZ<F>), st, fin_V, fin_R, g.H, fin_n, m_log,
                   g.c, mode, win, ticket, d_out_bytes, reinterpret_cast<XYZZ<F>*>(d_out_xyzz));
        } else {
            size_t smem = (size_t)BMPC_FINAL_THREADS * sizeof(XYZZ<F>);
            CK(cudaFuncSetAttribute(msm_final_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (groups) {
                size_t fsmem = (size_t)BMPC_FOLD_GROUP * sizeof(XYZZ<F>);
                dim3 fgrid(groups, g.H);
                LAUNCH(ctx, msm_fold_kernel<F>, fgrid, BMPC_FOLD_GROUP, fsmem, st, (const XYZZ<F>*)blk_V,
                       (const XYZZ<F>*)blk_R, p.nblk, m_log, grp_V, grp_R);
                for (uint32_t q = BMPC_FOLD_GROUP; q > 1; q >>= 1) m_log++;
                fin_V = grp_V; fin_R = grp_R; fin_n = groups;
            }
            LAUNCH(ctx, msm_final_kernel<F>, g.H, BMPC_FINAL_THREADS, smem, st, fin_V, fin_R,
                   g.H, fin_n, m_log, g.c, mode, win, ticket, d_out_bytes, reinterpret_cast<XYZZ<F>*>(d_out_xyzz));
        }
    }
    return BMPC_OK;
}

template <class F>
int GroupOps<F>::sum_partials(bmpc_ctx* ctx, const void* d_parts, uint32_t count, uint8_t* d_out_bytes,
                              cudaStream_t st) {
    LAUNCH(ctx, msm_sum_partials_kernel<F>, 1, 32, 0, st, reinterpret_cast<const XYZZ<F>*>(d_parts), count, d_out_bytes);
    return BMPC_OK;
}

template <class F>
int GroupOps<F>::decode(bmpc_ctx* ctx, const uint8_t* d_raw, size_t stride, size_t n, void* d_points,
                        cudaStream_t st) {
    if (!n) return BMPC_OK;
    LAUNCH(ctx, decode_uncompressed_kernel<F>, (uint32_t)((n + 127) / 128), 128, 0, st, d_raw, stride, n,
           reinterpret_cast<Affine<F>*>(d_points));
    return BMPC_OK;
}

template <class F>
int GroupOps<F>::validate_decode(bmpc_ctx* ctx, const uint8_t* d_raw, size_t stride, size_t n, int checked,
                                 int reject_identity, void* d_points, uint32_t* d_err, cudaStream_t st) {
    if (!n) return BMPC_OK;
    LAUNCH(ctx, validate_decode_kernel<F>, (uint32_t)((n + 63) / 64), 64, 0, st, d_raw, stride, n, checked,
           reject_identity, reinterpret_cast<Affine<F>*>(d_points), d_err);
    return BMPC_OK;
}

template <class F>
int GroupOps<F>::encode(bmpc_ctx* ctx, const void* d_points, size_t n, uint8_t* d_out, cudaStream_t st) {
    if (!n) return BMPC_OK;
    LAUNCH(ctx, encode_uncompressed_kernel<F>, (uint32_t)((n + 127) / 128), 128, 0, st,
           reinterpret_cast<const Affine<F>*>(d_points), n, d_out);
    return BMPC_OK;
}

template <class F>
int GroupOps<F>::inf_bitmap(bmpc_ctx* ctx, const void* d_points, size_t n, uint32_t* d_bitmap, cudaStream_t st) {
    if (!n) return BMPC_OK;
    LAUNCH(ctx, inf_bitmap_kernel<F>, (uint32_t)((n + 127) / 128), 128, 0, st,
           reinterpret_cast<const Affine<F>*>(d_points), n, d_bitmap);
    return BMPC_OK;
}

template <class F>
int GroupOps<F>::batch_mul(bmpc_ctx* ctx, const void* d_in, const uint32_t* d_scalars, int per_element,
                           size_t n, void* d_out, cudaStream_t st) {
    if (!n) return BMPC_OK;
    LAUNCH(ctx, batch_scalar_mul_kernel<F>, (uint32_t)((n + 127) / 128), 128, 0, st,
           reinterpret_cast<const Affine<F>*>(d_in), d_scalars, per_element, n, reinterpret_cast<Affine<F>*>(d_out));
    return BMPC_OK;
}

template <class F>
int GroupOps<F>::list_mul_matrix(bmpc_ctx* ctx, const void* d_list, const uint32_t* d_row_ptr, const uint32_t* d_col,
                                 const uint32_t* d_coeffs, size_t live_rows, size_t n_out, void* d_out,
                                 cudaStream_t st) {
    if (!n_out) return BMPC_OK;
    LAUNCH(ctx, list_mul_matrix_kernel<F>, (uint32_t)((n_out + 127) / 128), 128, 0, st,
           reinterpret_cast<const Affine<F>*>(d_list), d_row_ptr, d_col, d_coeffs, live_rows, n_out,
           reinterpret_cast<Affine<F>*>(d_out));
    return BMPC_OK;
}

template <class F>
int GroupOps<F>::fixed_base_mul(bmpc_ctx* ctx, const void* d_base, void* d_table, const uint32_t* d_scalars,
                                size_t n, void* d_out, cudaStream_t st) {
    LAUNCH(ctx, fixed_base_table_kernel<F>, 1, 32, 0, st, reinterpret_cast<const Affine<F>*>(d_base),
           reinterpret_cast<XYZZ<F>*>(d_table));
    if (n)
        LAUNCH(ctx, fixed_base_mul_kernel<F>, (uint32_t)((n + 127) / 128), 128, 0, st,
               reinterpret_cast<const XYZZ<F>*>(d_table), d_scalars, n, reinterpret_cast<Affine<F>*>(d_out));
    return BMPC_OK;
}

template <class F>
int GroupOps<F>::precompute_tables(bmpc_ctx* ctx, void* d_tables, size_t n, uint32_t c, uint32_t W,
                                   cudaStream_t st) {
    if (!n || W <= 1) return BMPC_OK;
    if (W > BMPC_MAX_TABLES) return BMPC_ERR_INVALID;
    LAUNCH(ctx, msm_precompute_kernel<F>, (uint32_t)((n + 63) / 64), 64, 0, st,
           reinterpret_cast<Affine<F>*>(d_tables), n, c, W);
    return BMPC_OK;
}

}  // namespace bmpc
