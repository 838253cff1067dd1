// Host-side launchers for the curve-typed kernels, instantiated once per group
// (group_g1.cu: F = Fp, group_g2.cu: F = Fp2).
#pragma once
#include <cstdlib>

#include "internal.h"
#include "group_kernels.cuh"

namespace bmpc {

template <class F>
size_t GroupOps<F>::curve_bytes(const MsmPlan& p) {
    size_t b = ws_need(p.max_tasks, sizeof(XYZZ<F>)) + ws_need((size_t)p.red_H * p.nblk, sizeof(XYZZ<F>)) + 1024;
    if (p.use2d)
        b += ws_need((size_t)p.g.H * 2 * p.NT, sizeof(XYZZ<F>)) + ws_need((size_t)p.red_H * p.Bm, sizeof(XYZZ<F>));
    return b;
}

template <class F>
int GroupOps<F>::msm_finish(bmpc_ctx* ctx, const MsmPlan& p, const bmpc_bases* bases, const MsmSorted& s,
                            int mode, uint8_t* d_out_bytes, void* d_out_xyzz, cudaStream_t st) {
    const MsmGeom& g = p.g;
    XYZZ<F>* partials = ws_take<XYZZ<F>>(ctx, p.max_tasks);
    XYZZ<F>* blk_out = ws_take<XYZZ<F>>(ctx, (size_t)p.red_H * p.nblk);
    XYZZ<F>* rc_part = nullptr;
    XYZZ<F>* rc_sums = nullptr;
    if (p.use2d) {
        rc_part = ws_take<XYZZ<F>>(ctx, (size_t)g.H * 2 * p.NT);
        rc_sums = ws_take<XYZZ<F>>(ctx, (size_t)p.red_H * p.Bm);
    }
    if (!partials || !blk_out || (p.use2d && (!rc_part || !rc_sums))) {
        ctx->err = "msm workspace carve failed (curve)";
        return BMPC_ERR_INVALID;
    }
    const Affine<F>* pts = reinterpret_cast<const Affine<F>*>(bases->d_points);
    uint32_t ablocks = (uint32_t)((p.max_tasks + 127) / 128);
    {
        ProfScope ps(ctx, BMPC_PROF_MSM_ACCUMULATE, st);
        // BMPC_ACC_COMPACT=1 selects the variant whose field products are calls (smaller code)
        static const bool compact = getenv("BMPC_ACC_COMPACT") && atoi(getenv("BMPC_ACC_COMPACT")) != 0;
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
                                (int)(128 * sizeof(XYZZ<F>))));
        dim3 grid(p.nblk, p.red_H);
        if (p.use2d) {
            dim3 g1((2 * p.NT + 127) / 128, g.H), g2((2 * p.Bm + 127) / 128, g.H);
            LAUNCH(ctx, msm_rowcol_kernel<F>, g1, 128, 0, st, (const XYZZ<F>*)partials, s.toff, g.B, p.logC, p.S2, rc_part);
            LAUNCH(ctx, msm_rowcol_fold_kernel<F>, g2, 128, 0, st, (const XYZZ<F>*)rc_part, g.B, p.logC, p.S2, p.Bm, rc_sums);
            LAUNCH(ctx, msm_reduce_kernel<F>, grid, p.rblock, smem, st, (const XYZZ<F>*)rc_sums, (const uint32_t*)nullptr,
                   p.red_B, p.S, blk_out);
        } else {
            LAUNCH(ctx, msm_reduce_kernel<F>, grid, p.rblock, smem, st, (const XYZZ<F>*)partials, s.toff, g.B, p.S, blk_out);
        }
    }
    {
        size_t smem = (size_t)(BMPC_FINAL_THREADS + p.red_H) * sizeof(XYZZ<F>);
        CK(cudaFuncSetAttribute(msm_final_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // doublings before adding an even / odd set (see msm_final_kernel)
        uint32_t into_even = p.use2d ? p.logC : g.c, into_odd = p.use2d ? g.c - p.logC : g.c;
        LAUNCH(ctx, msm_final_kernel<F>, 1, BMPC_FINAL_THREADS, smem, st, blk_out, p.red_H, p.nblk, into_even,
               into_odd, mode, d_out_bytes, reinterpret_cast<XYZZ<F>*>(d_out_xyzz));
    }
    return BMPC_OK;
}

template <class F>
int GroupOps<F>::sum_partials(bmpc_ctx* ctx, const void* d_parts, uint32_t count, uint8_t* d_out_bytes,
                              cudaStream_t st) {
    LAUNCH(ctx, msm_sum_partials_kernel<F>, 1, 1, 0, st, reinterpret_cast<const XYZZ<F>*>(d_parts), count, d_out_bytes);
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
