// Host side of the EvaluationDomain transforms: per-size constant/twiddle tables, pass
// planning, the fused H-polynomial pipeline.  Kernels are in ntt.cuh.
// Reference: src/domain.rs:47-189,261-372; src/groth16/prover.rs:210-231.
#include <cstdlib>
#include <cstring>

#include "internal.h"
#include "ntt.cuh"

namespace bmpc {

static int get_domain(bmpc_ctx* ctx, uint32_t logm, cudaStream_t st, DomainTables** out) {
    auto it = ctx->domains.find(logm);
    if (it == ctx->domains.end()) {
        DomainTables dt;
        CK(cudaMalloc(&dt.d_consts, 10 * sizeof(Fr)));
        LAUNCH(ctx, domain_consts_kernel, 1, 1, 0, st, logm, SMALL_LOG, dt.d_consts);
        CK(cudaMemcpyAsync(dt.h_consts, dt.d_consts, 10 * sizeof(Fr), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        it = ctx->domains.emplace(logm, dt).first;
        if (!ctx->tw_small[0]) {
            uint32_t cnt = 1u << (SMALL_LOG - 1);
            for (int dir = 0; dir < 2; dir++) {
                CK(cudaMalloc(&ctx->tw_small[dir], cnt * sizeof(Fr)));
                LAUNCH(ctx, pow_table_kernel, (cnt + 127) / 128, 128, 0, st,
                       (const Fr*)(it->second.d_consts + 8 + dir), (const Fr*)nullptr, (uint64_t)1, cnt, 0,
                       ctx->tw_small[dir]);
            }
        }
    }
    *out = &it->second;
    return BMPC_OK;
}

static int get_table(bmpc_ctx* ctx, DomainTables* dt, uint32_t logm, TableKind kind, cudaStream_t st,
                     PowTable* out) {
    DevTable& t = dt->t[kind];
    if (!t.ready) {
        int base_idx = 0, fold_idx = -1, canon = 0;
        switch (kind) {
            case K_TW_FWD: base_idx = 0; break;
            case K_TW_INV: base_idx = 1; break;
            case K_G: base_idx = 3; break;
            case K_G_MINV: base_idx = 3; fold_idx = 2; break;
            case K_GINV_MINV: base_idx = 4; fold_idx = 2; break;
            case K_GINV_MINV_ZINV_CANON: base_idx = 4; fold_idx = 6; canon = 1; break;
            default: return BMPC_ERR_INVALID;
        }
        t.lo_bits = logm <= 10 ? logm : (logm + 1) / 2;
        t.hi_n = 1u << (logm - t.lo_bits);
        uint32_t lo_n = 1u << t.lo_bits;
        CK(cudaMalloc(&t.lo, (size_t)lo_n * sizeof(Fr)));
        CK(cudaMalloc(&t.hi, (size_t)t.hi_n * sizeof(Fr)));
        const Fr* fold = fold_idx >= 0 ? dt->d_consts + fold_idx : nullptr;
        LAUNCH(ctx, pow_table_kernel, (lo_n + 127) / 128, 128, 0, st, (const Fr*)(dt->d_consts + base_idx), fold,
               (uint64_t)1, lo_n, canon, t.lo);
        LAUNCH(ctx, pow_table_kernel, (t.hi_n + 127) / 128, 128, 0, st, (const Fr*)(dt->d_consts + base_idx),
               (const Fr*)nullptr, (uint64_t)1 << t.lo_bits, t.hi_n, 0, t.hi);
        // 2-level tables cost one extra product per lookup; up to 2^24 the full table is expanded
        // once (32 B per coefficient in HBM, read once per pass that uses it)
        if (t.hi_n > 1 && logm <= DIRECT_TABLE_MAX_LOG && !ctx->tune.ntt_no_direct) {
            uint32_t cnt = 1u << logm;
            CK(cudaMalloc(&t.direct, (size_t)cnt * sizeof(Fr)));
            PowTable two{t.hi, t.lo, t.lo_bits, t.hi_n, nullptr};
            LAUNCH(ctx, pow_expand_kernel, (cnt + 255) / 256, 256, 0, st, two, cnt, t.direct);
        }
        t.ready = true;
    }
    out->hi = t.hi;
    out->lo = t.lo;
    out->lo_bits = t.lo_bits;
    out->hi_n = t.hi_n;
    out->direct = t.direct;
    return BMPC_OK;
}

void ntt_free_tables(bmpc_ctx* ctx) {
    for (auto& kv : ctx->domains) {
        cudaFree(kv.second.d_consts);
        for (int k = 0; k < K_COUNT; k++) {
            if (kv.second.t[k].hi) cudaFree(kv.second.t[k].hi);
            if (kv.second.t[k].lo) cudaFree(kv.second.t[k].lo);
            if (kv.second.t[k].direct) cudaFree(kv.second.t[k].direct);
        }
    }
    ctx->domains.clear();
    for (int d = 0; d < 2; d++) {
        if (ctx->tw_small[d]) cudaFree(ctx->tw_small[d]);
        ctx->tw_small[d] = nullptr;
    }
}

struct NttScale {
    int pre_mode = SCALE_NONE;
    TableKind pre_kind = K_G;
    int post_mode = SCALE_NONE;
    TableKind post_kind = K_GINV_MINV;
    int post_const_idx = 2;
};

// In-place transform of data[0 .. 2^logm); tmp1/tmp2: scratch buffers of the same size.
// `batch` > 1: that many independent transforms laid out back to back (data and scratch hold
// batch * 2^logm coefficients).
static int ntt_run(bmpc_ctx* ctx, Fr* data, Fr* tmp1, Fr* tmp2, uint32_t logm, bool inverse,
                   const NttScale& sc, cudaStream_t st, uint32_t batch = 1) {
    DomainTables* dt;
    int rc = get_domain(ctx, logm, st, &dt);
    if (rc) return rc;
    NttPassArgs A;
    memset(&A, 0, sizeof(A));
    A.logn = logm;
    A.batch_stride = (size_t)1 << logm;
    A.small_log = SMALL_LOG;
    A.tw_small = ctx->tw_small[inverse ? 1 : 0];
    PowTable pre{}, post{};
    if (sc.pre_mode == SCALE_POW) {
        rc = get_table(ctx, dt, logm, sc.pre_kind, st, &pre);
        if (rc) return rc;
    }
    if (sc.post_mode == SCALE_POW) {
        rc = get_table(ctx, dt, logm, sc.post_kind, st, &post);
        if (rc) return rc;
    }
    if (logm == 0) {
        A.in = data; A.out = data;
        A.pre_mode = sc.pre_mode; A.pre = pre;
        A.post_mode = sc.post_mode; A.post = post;
        A.post_const = dt->h_consts[sc.post_const_idx];
        if (batch != 1) return BMPC_ERR_INVALID;
        LAUNCH(ctx, ntt_trivial_kernel, 1, 1, 0, st, A);
        return BMPC_OK;
    }
    rc = get_table(ctx, dt, logm, inverse ? K_TW_INV : K_TW_FWD, st, &A.tw);
    if (rc) return rc;
    // 2^8 per pass with 4 sub-transforms per block measured fastest on B200 (2^24: 4.99 ms vs 5.58 ms
    // for two 2^12 passes, whose single-sub-transform layout has shared-memory bank conflicts in the
    // late stages); larger radices stay available through the tuning knob.
    uint32_t maxdeg = ctx->tune_maxdeg ? (uint32_t)ctx->tune_maxdeg : 8u;
    if (maxdeg > SMALL_LOG) maxdeg = SMALL_LOG;
    if (maxdeg < 1) maxdeg = 1;
    uint32_t npass = (logm + maxdeg - 1) / maxdeg;
    uint32_t plog = 0;
    const Fr* src = data;
    for (uint32_t j = 0; j < npass; j++) {
        uint32_t deg = logm / npass + (j < logm % npass ? 1u : 0u);
        Fr* dst;
        if (npass == 1) dst = tmp1;
        else if (j == npass - 1) dst = data;
        else dst = (j & 1) ? tmp2 : tmp1;
        uint32_t tlog = logm - deg;
        // T consecutive i per block while the tile stays within 32 KB; big radices run T = 1
        uint32_t tile_log = deg >= 10 ? 0 : (10 - deg > 2 ? 2 : 10 - deg);
        if (tile_log > tlog) tile_log = tlog;
        A.in = src; A.out = dst;
        A.deg = deg; A.plog = plog; A.tile_log = tile_log;
        A.pre_mode = (j == 0) ? sc.pre_mode : SCALE_NONE;
        A.pre = pre;
        A.post_mode = (j == npass - 1) ? sc.post_mode : SCALE_NONE;
        A.post = post;
        if (A.post_mode == SCALE_CONST) A.post_const = dt->h_consts[sc.post_const_idx];
        uint32_t threads = 1u << (tile_log + deg - 1);
        if (threads > 1024) threads = 1024;
        uint32_t blocks = 1u << (tlog - tile_log);
        size_t smem = ((size_t)1 << (tile_log + deg)) * sizeof(Fr);
        if (smem > 48 * 1024)
            CK(cudaFuncSetAttribute(ntt_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        {
            ProfScope ps(ctx, BMPC_PROF_NTT_PASS, st);
            LAUNCH(ctx, ntt_pass_kernel, dim3(blocks, batch), threads, smem, st, A);
        }
        src = dst;
        plog += deg;
    }
    if (npass == 1)
        CK(cudaMemcpyAsync(data, tmp1, ((size_t)batch << logm) * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
    return BMPC_OK;
}

// ---- pieces of the distributed four-step transform (bellman_mpc_b200/dist.py: DistributedDomain)
int ntt_batch_dev_locked(bmpc_ctx* ctx, Fr* d, uint32_t logn, uint32_t batch, bool inverse, cudaStream_t st) {
    if (logn >= 32 || logn == 0 || batch == 0 || batch > 65535) return BMPC_ERR_INVALID;
    size_t total = (size_t)batch << logn;
    int rc = ws_reserve(ctx, 2 * ws_need(total, sizeof(Fr)));
    if (rc) return rc;
    Fr* t1 = ws_take<Fr>(ctx, total);
    Fr* t2 = ws_take<Fr>(ctx, total);
    NttScale plain;
    return ntt_run(ctx, d, t1, t2, logn, inverse, plain, st, batch);
}

int fr_swap01(bmpc_ctx* ctx, const Fr* in, Fr* out, uint32_t d0, uint32_t d1, uint32_t d2, cudaStream_t st) {
    size_t total = (size_t)d0 * d1 * d2;
    if (!total) return BMPC_OK;
    if (d2 >= 8) {
        LAUNCH(ctx, fr_swap01_rows_kernel, (uint32_t)((total + 255) / 256), 256, 0, st, in, out, d0, d1, d2);
    } else {
        dim3 grid((d1 + BMPC_TR_TILE - 1) / BMPC_TR_TILE, (d0 + BMPC_TR_TILE - 1) / BMPC_TR_TILE, 1);
        if (grid.y > 65535) return BMPC_ERR_INVALID;
        LAUNCH(ctx, fr_swap01_kernel, grid, BMPC_TR_TILE * BMPC_TR_TILE, 0, st, in, out, d0, d1, d2);
    }
    return BMPC_OK;
}

int fr_fourstep_twiddle(bmpc_ctx* ctx, Fr* d, uint32_t rows, uint32_t cols, uint32_t row0, uint32_t logm,
                        bool inverse, cudaStream_t st) {
    DomainTables* dt;
    int rc = get_domain(ctx, logm, st, &dt);
    if (rc) return rc;
    PowTable tw;
    rc = get_table(ctx, dt, logm, inverse ? K_TW_INV : K_TW_FWD, st, &tw);
    if (rc) return rc;
    size_t total = (size_t)rows * cols;
    if (!total) return BMPC_OK;
    LAUNCH(ctx, fr_fourstep_twiddle_kernel, (uint32_t)((total + 255) / 256), 256, 0, st, d, rows, cols, row0, logm, tw);
    return BMPC_OK;
}

// which: 0 = g^i (coset shift), 1 = g^-i / m (inverse coset), 2 = 1 / m (inverse)
int fr_scale_pow(bmpc_ctx* ctx, Fr* d, size_t n, uint32_t first, uint32_t logm, int which, cudaStream_t st) {
    DomainTables* dt;
    int rc = get_domain(ctx, logm, st, &dt);
    if (rc) return rc;
    PowTable T{};
    if (which == 0 || which == 1) {
        rc = get_table(ctx, dt, logm, which == 0 ? K_G : K_GINV_MINV, st, &T);
        if (rc) return rc;
    } else if (which != 2) {
        return BMPC_ERR_INVALID;
    }
    if (!n) return BMPC_OK;
    LAUNCH(ctx, fr_scale_pow_kernel, (uint32_t)((n + 255) / 256), 256, 0, st, d, n, first, T, (const Fr*)(dt->d_consts + 2));
    return BMPC_OK;
}

int ntt_dev_locked(bmpc_ctx* ctx, Fr* d, uint32_t logm, int op, cudaStream_t st) {
    if (logm >= 32) return BMPC_ERR_DEGREE_TOO_LARGE;
    bool inverse = false;
    NttScale sc;
    switch (op) {
        case BMPC_FFT: break;
        case BMPC_IFFT: inverse = true; sc.post_mode = SCALE_CONST; sc.post_const_idx = 2; break;
        case BMPC_COSET_FFT: sc.pre_mode = SCALE_POW; sc.pre_kind = K_G; break;
        case BMPC_ICOSET_FFT: inverse = true; sc.post_mode = SCALE_POW; sc.post_kind = K_GINV_MINV; break;
        default: return BMPC_ERR_INVALID;
    }
    size_t m = (size_t)1 << logm;
    int rc = ws_reserve(ctx, 2 * ws_need(m, sizeof(Fr)));
    if (rc) return rc;
    Fr* t1 = ws_take<Fr>(ctx, m);
    Fr* t2 = ws_take<Fr>(ctx, m);
    return ntt_run(ctx, d, t1, t2, logm, inverse, sc, st);
}

// The H pipeline of prover.rs:214-226 in its two halves (so that several GPUs can share it: one
// vector each, then one device combines -- bellman_mpc_b200/dist.py):
//   h_coset_evals_locked:      P <- iNTT(P) (unscaled); P <- NTT(P_i * g^i / m): the evaluations of the
//                              polynomial through `P` on the coset g <omega>  (ifft + coset_fft)
//   h_from_coset_evals_locked: a <- a*b - c ; a <- iNTT(a) ; a_i <- a_i * g^-i / (m Z(g)), canonical
//                              (mul_assign, sub_assign, divide_by_z_on_coset, icoset_fft, to_le_bits)
// with the O(m) sweeps folded into the transforms.
int h_coset_evals_locked(bmpc_ctx* ctx, Fr* p, uint32_t logm, Fr* t1, Fr* t2, cudaStream_t st) {
    NttScale inv_plain;
    NttScale coset_fused;
    coset_fused.pre_mode = SCALE_POW;
    coset_fused.pre_kind = K_G_MINV;
    int rc = ntt_run(ctx, p, t1, t2, logm, true, inv_plain, st);
    if (rc) return rc;
    return ntt_run(ctx, p, t1, t2, logm, false, coset_fused, st);
}

int h_from_coset_evals_locked(bmpc_ctx* ctx, Fr* a, const Fr* b, const Fr* c, uint32_t logm, Fr* t1, Fr* t2,
                              cudaStream_t st) {
    size_t m = (size_t)1 << logm;
    LAUNCH(ctx, fr_mul_sub_kernel, (uint32_t)((m + 255) / 256), 256, 0, st, a, b, c, m);
    NttScale fin;
    fin.post_mode = SCALE_POW;
    fin.post_kind = K_GINV_MINV_ZINV_CANON;
    return ntt_run(ctx, a, t1, t2, logm, true, fin, st);
}

int h_coefficients_locked(bmpc_ctx* ctx, Fr* a, Fr* b, Fr* c, uint32_t logm, Fr* t1, Fr* t2,
                          cudaStream_t st) {
    Fr* polys[3] = {a, b, c};
    for (int k = 0; k < 3; k++) {
        int rc = h_coset_evals_locked(ctx, polys[k], logm, t1, t2, st);
        if (rc) return rc;
    }
    return h_from_coset_evals_locked(ctx, a, b, c, logm, t1, t2, st);
}

int fr_pointwise(bmpc_ctx* ctx, int what, Fr* a, const Fr* b, size_t n, cudaStream_t st) {
    if (!n) return BMPC_OK;
    uint32_t blocks = (uint32_t)((n + 255) / 256);
    if (what == 0) LAUNCH(ctx, fr_mul_assign_kernel, blocks, 256, 0, st, a, b, n);
    else if (what == 1) LAUNCH(ctx, fr_sub_assign_kernel, blocks, 256, 0, st, a, b, n);
    else LAUNCH(ctx, fr_to_canonical_kernel, blocks, 256, 0, st, a, n);
    return BMPC_OK;
}

int fr_scale_zinv(bmpc_ctx* ctx, Fr* a, size_t m, uint32_t logm, cudaStream_t st) {
    DomainTables* dt;
    int rc = get_domain(ctx, logm, st, &dt);
    if (rc) return rc;
    LAUNCH(ctx, fr_scale_kernel, (uint32_t)((m + 255) / 256), 256, 0, st, a, (const Fr*)(dt->d_consts + 5), m);
    return BMPC_OK;
}

int fr_distribute_powers(bmpc_ctx* ctx, Fr* a, size_t m, const Fr* d_g, cudaStream_t st) {
    size_t threads = (m + 7) / 8;
    LAUNCH(ctx, fr_distribute_powers_kernel, (uint32_t)((threads + 127) / 128), 128, 0, st, a, d_g, m);
    return BMPC_OK;
}

int fr_eval_z(bmpc_ctx* ctx, const Fr* d_tau, uint32_t logm, Fr* d_out, cudaStream_t st) {
    LAUNCH(ctx, fr_z_kernel, 1, 1, 0, st, d_tau, logm, d_out);
    return BMPC_OK;
}

}  // namespace bmpc
