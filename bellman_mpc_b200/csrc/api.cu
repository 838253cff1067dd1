// libbellman_b200.so -- the C ABI declared in include/bellman_b200.h.
// This file owns contexts, device memory and the reference's error semantics; kernels and
// their launch geometry live in ntt.cu / msm_sort.cu / group_g{1,2}.cu / prove.cu.  No field
// or group arithmetic runs on the host: every constant is computed by a device kernel.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

#include "internal.h"

using namespace bmpc;

namespace bmpc {

int flags_to_status(uint32_t flags) {
    // SURVEY 8a'/5: EOF fails every window; an identity only the windows that consume it.
    if (flags & MSM_FLAG_EOF) return (flags & MSM_FLAG_IDENT_TOP) ? BMPC_ERR_UNEXPECTED_IDENTITY : BMPC_ERR_UNEXPECTED_EOF;
    if (flags & MSM_FLAG_IDENT_ANY) return BMPC_ERR_UNEXPECTED_IDENTITY;
    return BMPC_OK;
}

void write_identity(int group, uint8_t* out) {
    size_t nb = group == BMPC_G1 ? 96 : 192;
    memset(out, 0, nb);
    out[0] = 0x40;
}

// n_ref: the length of the WHOLE exponent vector when this call is one shard of it (the reference
// derives its window size -- and with it which error wins -- from that length, multiexp.rs:267-271);
// 0 = this call is the whole multiexp.
int multiexp_enqueue(bmpc_ctx* ctx, const bmpc_bases* bases, size_t base_offset, const uint64_t* d_scalars,
                     size_t n, const uint64_t* d_density, size_t density_len, uint8_t* out, void* d_partial,
                     cudaStream_t st, uint8_t* h_dst, MsmPending* pend, size_t n_ref) {
    *pend = MsmPending();
    if (!bases || (!out && !d_partial)) return BMPC_ERR_INVALID;
    if (d_density && density_len != n) return BMPC_ERR_LENGTH_MISMATCH;  // multiexp.rs:273-278
    if (bases->n >= ((size_t)1 << 31) || base_offset >= ((size_t)1 << 31) || n >= ((size_t)1 << 31))
        return BMPC_ERR_INVALID;
    const size_t out_bytes = bases->group == BMPC_G1 ? 96 : 192;
    if (n == 0) {  // SURVEY 8a'/8: identity, no error
        if (d_partial) CK(cudaMemsetAsync(d_partial, 0, bmpc_partial_bytes(bases->group), st));
        if (out) write_identity(bases->group, out);
        return BMPC_OK;
    }
    int rc;
    uint32_t* d_flags = reinterpret_cast<uint32_t*>(ctx->d_stage);
    uint8_t* d_bytes = ctx->d_stage + 64;
    int mode = d_partial ? 1 : 0;
    const size_t dense_hint = ctx->dense_hint;
    ctx->dense_hint = 0;
    MsmPlan p = msm_make_plan(ctx, bases, n, d_density != nullptr, n_ref, dense_hint);
    // histogram, offsets, cursors and task descriptors are 32-bit over the n * W (position, window)
    // pairs: refuse what would overflow them instead of corrupting the sort
    if (p.max_pairs >= ((size_t)1 << 32) - 1) {
        ctx->err = "multiexp too large for one call: n * windows >= 2^32 (split the exponent range)";
        return BMPC_ERR_INVALID;
    }
    if (bases->group == BMPC_G1) GroupOps<Fp>::plan_affine(ctx, p);
    else GroupOps<Fp2>::plan_affine(ctx, p);
    size_t curve_bytes = bases->group == BMPC_G1 ? GroupOps<Fp>::curve_bytes(p) : GroupOps<Fp2>::curve_bytes(p);
    rc = ws_reserve(ctx, p.sort_bytes + curve_bytes);
    if (rc) return rc;
    MsmSorted sorted;
    rc = msm_sort_run(ctx, p, bases, base_offset, (const uint32_t*)d_scalars, n, (const uint32_t*)d_density,
                      d_flags, &sorted, st);
    if (rc) return rc;
    if (bases->group == BMPC_G1)
        rc = GroupOps<Fp>::msm_finish(ctx, p, bases, sorted, mode, d_bytes, d_partial, st);
    else
        rc = GroupOps<Fp2>::msm_finish(ctx, p, bases, sorted, mode, d_bytes, d_partial, st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h_dst, ctx->d_stage, 64 + out_bytes, cudaMemcpyDeviceToHost, st));
    pend->st = st;
    pend->h = h_dst;
    pend->out_bytes = out_bytes;
    pend->out = out;
    return BMPC_OK;
}

// after the stream was synchronised
int multiexp_collect(const MsmPending& pend, uint32_t* flags_out) {
    if (flags_out) *flags_out = 0;
    if (!pend.st) return BMPC_OK;
    uint32_t flags = *reinterpret_cast<const uint32_t*>(pend.h);
    if (flags_out) *flags_out = flags;
    int status = flags_to_status(flags);
    if (status == BMPC_OK && pend.out) memcpy(pend.out, pend.h + 64, pend.out_bytes);
    return status;
}

int multiexp_dev_locked(bmpc_ctx* ctx, const bmpc_bases* bases, size_t base_offset,
                        const uint64_t* d_scalars, size_t n, const uint64_t* d_density,
                        size_t density_len, uint8_t* out, void* d_partial, cudaStream_t st,
                        size_t n_ref, uint32_t* flags_out) {
    MsmPending pend;
    int rc = multiexp_enqueue(ctx, bases, base_offset, d_scalars, n, d_density, density_len, out, d_partial, st,
                              ctx->h_stage, &pend, n_ref);
    if (rc) return rc;
    if (pend.st) CK(cudaStreamSynchronize(st));
    return multiexp_collect(pend, flags_out);
}

}  // namespace bmpc

namespace {

// a bmpc_bases under construction: device memory and the handle are released unless handed out
struct BasesGuard {
    bmpc_ctx* ctx;
    bmpc_bases* b = nullptr;
    explicit BasesGuard(bmpc_ctx* c) : ctx(c) {}
    bmpc_bases* release() { bmpc_bases* r = b; b = nullptr; return r; }
    ~BasesGuard() {
        if (!b) return;
        cudaStreamSynchronize(ctx->own_stream);
        if (b->d_points) cudaFree(b->d_points);
        if (b->d_inf) cudaFree(b->d_inf);
        delete b;
    }
};

int register_points(bmpc_ctx* ctx, bmpc_bases* b, cudaStream_t st) {
    size_t nw = (b->n + 31) / 32 + 1;
    CK(cudaMalloc(&b->d_inf, nw * 4));
    CK(cudaMemsetAsync(b->d_inf, 0, nw * 4, st));
    return b->group == BMPC_G1 ? GroupOps<Fp>::inf_bitmap(ctx, b->d_points, b->n, b->d_inf, st)
                               : GroupOps<Fp2>::inf_bitmap(ctx, b->d_points, b->n, b->d_inf, st);
}

}  // namespace

void bmpc_tuning::load() {
    auto geti = [](const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; };
    acc_pairs = geti("BMPC_ACC_PAIRS", 0);
    const char* me = getenv("BMPC_PAIR_MIN_ENTRIES");
    pair_min_entries = me ? (size_t)atoll(me) : ((size_t)1 << 22);
    pair_k = geti("BMPC_PAIR_K", 0);
    acc_affine = geti("BMPC_ACC_AFFINE", -1);
    aff_blockdim = geti("BMPC_AFF_BLOCKDIM", 0);
    aff_ksel = geti("BMPC_AFF_KSEL", 0);
    aff_minb = geti("BMPC_AFF_MINB", 0);
    aff_gmax = geti("BMPC_AFF_GMAX", 0);
    aff_waves = geti("BMPC_AFF_WAVES", 0);
    aff_force_g = geti("BMPC_AFF_FORCE_G", 0);
    aff_whole_waves = geti("BMPC_AFF_WHOLE_WAVES", 1);
    acc_compact = geti("BMPC_ACC_COMPACT", 0);
    subwindows = geti("BMPC_MSM_SUBWINDOWS", 1);
    reduce_block = geti("BMPC_REDUCE_BLOCK", 0);
    reduce_slog_add = geti("BMPC_REDUCE_SLOG_ADD", 0);
    ntt_no_direct = geti("BMPC_NTT_NO_DIRECT", 0);
    proof_slots = geti("BMPC_PROOF_SLOTS", 0);
    tail_quad = geti("BMPC_TAIL_QUAD", 1);
    sort_radix = geti("BMPC_SORT_RADIX", -1);
    rs_chunk_log = geti("BMPC_RS_CHUNK_LOG", 0);
}

// =============================================================================== C ABI
extern "C" {

int bmpc_ctx_create(int device, bmpc_ctx** out) {
    if (!out) return BMPC_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return BMPC_ERR_CUDA;
    bmpc_ctx* ctx = new bmpc_ctx();
    ctx->device = device;
    DeviceGuard dg(device);
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMallocHost(&ctx->h_stage, 4096) != cudaSuccess ||
        cudaMalloc(&ctx->d_stage, 4096) != cudaSuccess) {
        delete ctx;
        return BMPC_ERR_CUDA;
    }
    ctx->tune.load();
    const char* ec = getenv("BMPC_MSM_WINDOW");
    if (ec) ctx->tune_c = atoi(ec);
    const char* ed = getenv("BMPC_NTT_MAXDEG");
    if (ed) ctx->tune_maxdeg = atoi(ed);
    *out = ctx;
    return BMPC_OK;
}

void bmpc_ctx_destroy(bmpc_ctx* ctx) {
    if (!ctx) return;
    for (bmpc_ctx* lane : ctx->lanes) bmpc_ctx_destroy(lane);
    ctx->lanes.clear();
    DeviceGuard dg(ctx->device);
    cudaDeviceSynchronize();
    ntt_free_tables(ctx);
    if (ctx->ws) cudaFree(ctx->ws);
    for (auto& sl : ctx->slots) {
        if (sl.ws) cudaFree(sl.ws);
        if (sl.d_stage) cudaFree(sl.d_stage);
        if (sl.stream) cudaStreamDestroy(sl.stream);
    }
    if (ctx->io) cudaFree(ctx->io);
    if (ctx->d_stage) cudaFree(ctx->d_stage);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->copy_done) cudaEventDestroy(ctx->copy_done);
    if (ctx->last_done) cudaEventDestroy(ctx->last_done);
    if (ctx->inputs_ready) cudaEventDestroy(ctx->inputs_ready);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char* bmpc_last_error(const bmpc_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int bmpc_ctx_set_tuning(bmpc_ctx* ctx, int msm_window_bits, int ntt_max_deg) {
    if (!ctx) return BMPC_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->tune_c = msm_window_bits;
    ctx->tune_maxdeg = ntt_max_deg;
    for (bmpc_ctx* lane : ctx->lanes) lane->tune_c = msm_window_bits;
    return BMPC_OK;
}

int bmpc_ctx_reload_env(bmpc_ctx* ctx) {
    if (!ctx) return BMPC_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->tune.load();
    for (bmpc_ctx* lane : ctx->lanes) lane->tune = ctx->tune;
    return BMPC_OK;
}

uint64_t bmpc_ctx_launch_count(const bmpc_ctx* ctx) {
    if (!ctx) return 0;
    uint64_t total = ctx->launches;
    for (const bmpc_ctx* lane : ctx->lanes) total += lane->launches;
    return total;
}

int bmpc_ctx_profile(bmpc_ctx* ctx, int enable) {
    if (!ctx) return BMPC_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->profile = enable != 0;
    return BMPC_OK;
}

int bmpc_ctx_profile_read(bmpc_ctx* ctx, int which, double* ms_total, uint64_t* launches) {
    if (!ctx || which < 0 || which >= BMPC_PROF_COUNT) return BMPC_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    CK(cudaDeviceSynchronize());
    for (auto& ev : ctx->prof_pending) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ev.e0, ev.e1) == cudaSuccess) {
            ctx->prof_ms[ev.id] += ms;
            ctx->prof_n[ev.id] += 1;
        }
        cudaEventDestroy(ev.e0);
        cudaEventDestroy(ev.e1);
    }
    ctx->prof_pending.clear();
    if (ms_total) *ms_total = ctx->prof_ms[which];
    if (launches) *launches = ctx->prof_n[which];
    ctx->prof_ms[which] = 0;
    ctx->prof_n[which] = 0;
    return BMPC_OK;
}

// ---------------------------------------------------------------------------- bases
int bmpc_bases_register(bmpc_ctx* ctx, int group, const void* points, size_t n, size_t stride,
                        int form, bmpc_bases** out) {
    if (!ctx || !out || (group != BMPC_G1 && group != BMPC_G2) || (!points && n)) return BMPC_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = ctx->own_stream;
    StreamScope ss(ctx, st);
    size_t pb = group == BMPC_G1 ? 96 : 192;
    if (stride == 0) stride = pb;
    if (stride < pb) return BMPC_ERR_INVALID;
    if (form != BMPC_FORM_MONT_XY && form != BMPC_FORM_UNCOMPRESSED_BE) return BMPC_ERR_INVALID;
    BasesGuard bg(ctx);
    bmpc_bases* b = bg.b = new bmpc_bases();
    b->group = group;
    b->n = n;
    CK(cudaMalloc(&b->d_points, (n ? n : 1) * pb));
    DevBuf raw;
    if (n) {
        if (form == BMPC_FORM_MONT_XY) {
            CK(cudaMemcpy2DAsync(b->d_points, pb, points, stride, pb, n, cudaMemcpyHostToDevice, st));
        } else {
            CK(cudaMalloc(&raw.p, n * stride));
            CK(cudaMemcpyAsync(raw.p, points, n * stride, cudaMemcpyHostToDevice, st));
            int rcd = group == BMPC_G1 ? GroupOps<Fp>::decode(ctx, raw.as<uint8_t>(), stride, n, b->d_points, st)
                                       : GroupOps<Fp2>::decode(ctx, raw.as<uint8_t>(), stride, n, b->d_points, st);
            if (rcd) return rcd;
        }
    }
    int rc = register_points(ctx, b, st);
    if (rc) return rc;
    CK(cudaStreamSynchronize(st));
    *out = bg.release();
    return BMPC_OK;
}

int bmpc_bases_register_dev(bmpc_ctx* ctx, int group, const void* d_points_mont, size_t n,
                            bmpc_bases** out, void* stream) {
    if (!ctx || !out || (group != BMPC_G1 && group != BMPC_G2)) return BMPC_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = pick_stream(ctx, stream);
    StreamScope ss(ctx, st);
    size_t pb = group == BMPC_G1 ? 96 : 192;
    BasesGuard bg(ctx);
    bmpc_bases* b = bg.b = new bmpc_bases();
    b->group = group;
    b->n = n;
    CK(cudaMalloc(&b->d_points, (n ? n : 1) * pb));
    if (n) CK(cudaMemcpyAsync(b->d_points, d_points_mont, n * pb, cudaMemcpyDeviceToDevice, st));
    int rc = register_points(ctx, b, st);
    if (rc) return rc;
    CK(cudaStreamSynchronize(st));
    *out = bg.release();
    return BMPC_OK;
}

int bmpc_bases_precompute(bmpc_ctx* ctx, bmpc_bases* b, int window_bits) {
    if (!ctx || !b) return BMPC_ERR_INVALID;
    if (b->tab_W || b->n == 0) return BMPC_OK;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = ctx->own_stream;
    StreamScope ss(ctx, st);
    uint32_t c = window_bits > 0 ? (uint32_t)window_bits : msm_table_window(b->n);
    if (c < 2) c = 2;
    if (c > 22) c = 22;
    uint32_t W = 255 / c + 1;
    if (W > 32 || (size_t)W * b->n >= ((size_t)1 << 31)) return BMPC_OK;  // not worth it / index range: stay table-free
    size_t pb = b->group == BMPC_G1 ? 96 : 192;
    void* d_tab;
    CK(cudaMalloc(&d_tab, (size_t)W * b->n * pb));
    CK(cudaMemcpyAsync(d_tab, b->d_points, b->n * pb, cudaMemcpyDeviceToDevice, st));
    int rc = b->group == BMPC_G1 ? GroupOps<Fp>::precompute_tables(ctx, d_tab, b->n, c, W, st)
                                 : GroupOps<Fp2>::precompute_tables(ctx, d_tab, b->n, c, W, st);
    if (rc) { cudaFree(d_tab); return rc; }
    CK(cudaStreamSynchronize(st));
    CK(cudaFree(b->d_points));
    b->d_points = d_tab;
    b->tab_c = c;
    b->tab_W = W;
    return BMPC_OK;
}

size_t bmpc_bases_len(const bmpc_bases* b) { return b ? b->n : 0; }
int bmpc_bases_group(const bmpc_bases* b) { return b ? b->group : 0; }
const void* bmpc_bases_dev_ptr(const bmpc_bases* b) { return b ? b->d_points : nullptr; }

int bmpc_bases_read(bmpc_ctx* ctx, const bmpc_bases* b, size_t start, size_t count, uint8_t* out) {
    if (!ctx || !b || !out || start + count > b->n) return BMPC_ERR_INVALID;
    if (!count) return BMPC_OK;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = ctx->own_stream;
    StreamScope ss(ctx, st);
    size_t pb = b->group == BMPC_G1 ? 96 : 192;
    DevBuf obuf;
    CK(cudaMalloc(&obuf.p, count * pb));
    uint8_t* d_out = obuf.as<uint8_t>();
    int rce = b->group == BMPC_G1
                  ? GroupOps<Fp>::encode(ctx, (const G1Affine*)b->d_points + start, count, d_out, st)
                  : GroupOps<Fp2>::encode(ctx, (const G2Affine*)b->d_points + start, count, d_out, st);
    if (rce) return rce;
    CK(cudaMemcpyAsync(out, d_out, count * pb, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return BMPC_OK;
}

void bmpc_bases_free(bmpc_ctx* ctx, bmpc_bases* b) {
    if (!b) return;
    if (ctx) {
        DeviceGuard dg(ctx->device);
        cudaDeviceSynchronize();
        if (b->d_points) cudaFree(b->d_points);
        if (b->d_inf) cudaFree(b->d_inf);
    }
    delete b;
}

// ------------------------------------------------------------------------- multiexp
size_t bmpc_partial_bytes(int group) { return group == BMPC_G1 ? sizeof(G1XYZZ) : sizeof(G2XYZZ); }

int bmpc_msm_geometry(bmpc_ctx* ctx, const bmpc_bases* bases, size_t n, uint32_t* window_bits,
                      uint32_t* windows, uint32_t* bucket_sets) {
    if (!ctx || !bases) return BMPC_ERR_INVALID;
    MsmPlan p = msm_make_plan(ctx, bases, n, false);
    if (window_bits) *window_bits = p.g.c;
    if (windows) *windows = p.g.W;
    if (bucket_sets) *bucket_sets = p.g.H;
    return BMPC_OK;
}

int bmpc_msm_accumulate_info(bmpc_ctx* ctx, const bmpc_bases* bases, size_t n, uint32_t info[8]) {
    if (!ctx || !bases || !info) return BMPC_ERR_INVALID;
    DeviceGuard dg(ctx->device);
    MsmPlan p = msm_make_plan(ctx, bases, n, false);
    if (bases->group == BMPC_G1) GroupOps<Fp>::plan_affine(ctx, p);
    else GroupOps<Fp2>::plan_affine(ctx, p);
    for (int i = 0; i < 8; i++) info[i] = 0;
    info[0] = p.pairs ? 2u : (p.affine ? 1u : 0u);
    info[1] = p.aff_G; info[2] = p.aff_K; info[3] = p.aff_blocks; info[4] = p.aff_block; info[5] = p.g.L;
    if (p.pairs) { info[1] = p.pair_R; info[2] = p.pair_kmax; info[3] = p.pair_blocks; info[4] = p.pair_block; }
    return BMPC_OK;
}

int bmpc_multiexp_dev(bmpc_ctx* ctx, const bmpc_bases* bases, size_t base_offset,
                      const uint64_t* d_scalars, size_t n, const uint64_t* d_density_words,
                      size_t density_len, uint8_t* out, void* stream) {
    if (!ctx || !out) return BMPC_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = pick_stream(ctx, stream);
    StreamScope ss(ctx, st);
    return multiexp_dev_locked(ctx, bases, base_offset, d_scalars, n, d_density_words, density_len,
                               out, nullptr, st);
}

int bmpc_multiexp_partial_dev(bmpc_ctx* ctx, const bmpc_bases* bases, size_t base_offset,
                              const uint64_t* d_scalars, size_t n, const uint64_t* d_density_words,
                              size_t density_len, void* d_partial_out, void* stream) {
    if (!ctx || !d_partial_out) return BMPC_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = pick_stream(ctx, stream);
    StreamScope ss(ctx, st);
    return multiexp_dev_locked(ctx, bases, base_offset, d_scalars, n, d_density_words, density_len,
                               nullptr, d_partial_out, st);
}

int bmpc_multiexp_shard_dev(bmpc_ctx* ctx, const bmpc_bases* bases, size_t base_offset,
                            const uint64_t* d_scalars, size_t n, const uint64_t* d_density_words,
                            size_t density_len, size_t n_total, void* d_partial_out, uint32_t* flags_out,
                            void* stream) {
    if (!ctx || !d_partial_out || !flags_out || n_total < n) return BMPC_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = pick_stream(ctx, stream);
    StreamScope ss(ctx, st);
    int rc = multiexp_dev_locked(ctx, bases, base_offset, d_scalars, n, d_density_words, density_len,
                                 nullptr, d_partial_out, st, n_total, flags_out);
    // the shard's own status is not the multiexp's: the caller ORs the flag words of all shards
    return (rc == BMPC_ERR_UNEXPECTED_EOF || rc == BMPC_ERR_UNEXPECTED_IDENTITY) ? BMPC_OK : rc;
}

// Enqueue-only form of the shard call: nothing is synchronised; the shard's RECORD -- its XYZZ partial
// followed by its raw flag word -- is left at d_record_out (bmpc_shard_record_bytes) in stream order, so
// the caller can all-gather the records of all ranks and fold them with ONE synchronisation
// (bmpc_fold_shard_records) instead of one per step.
int bmpc_multiexp_shard_enqueue_dev(bmpc_ctx* ctx, const bmpc_bases* bases, size_t base_offset,
                                    const uint64_t* d_scalars, size_t n, const uint64_t* d_density_words,
                                    size_t density_len, size_t n_total, void* d_record_out, void* stream) {
    if (!ctx || !bases || !d_record_out || n_total < n) return BMPC_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = pick_stream(ctx, stream);
    StreamScope ss(ctx, st);
    const size_t pb = bmpc_partial_bytes(bases->group);
    MsmPending pend;
    int rc = multiexp_enqueue(ctx, bases, base_offset, d_scalars, n, d_density_words, density_len, nullptr,
                              d_record_out, st, ctx->h_stage, &pend, n_total);
    if (rc) return rc;
    uint8_t* tail = reinterpret_cast<uint8_t*>(d_record_out) + pb;
    CK(cudaMemsetAsync(tail, 0, 16, st));
    if (pend.st) CK(cudaMemcpyAsync(tail, ctx->d_stage, 4, cudaMemcpyDeviceToDevice, st));   // the flag word
    return BMPC_OK;
}

size_t bmpc_shard_record_bytes(int group) { return bmpc_partial_bytes(group) + 16; }

namespace {
// records (partial + flag word, `stride` bytes apart) -> contiguous partials + OR of the flag words
__global__ void shard_records_unpack_kernel(const uint32_t* rec, uint32_t stride_words, uint32_t count,
                                            uint32_t pwords, uint32_t* contig, uint32_t* flags_or) {
    const uint32_t i = blockIdx.x;
    if (i >= count) return;
    const uint32_t* r = rec + (size_t)i * stride_words;
    for (uint32_t w = threadIdx.x; w < pwords; w += blockDim.x) contig[(size_t)i * pwords + w] = r[w];
    if (threadIdx.x == 0 && r[pwords]) atomicOr(flags_or, r[pwords]);
}
}  // namespace

// Folds `count` shard records (as left by bmpc_multiexp_shard_enqueue_dev, all-gathered by the caller;
// `stride` bytes apart) on `stream`: sum of the partials -> out (uncompressed affine), OR of the flag words
// -> *flags_or_out; the multiexp's status is bmpc_msm_flags_status(*flags_or_out), and `out` is only
// meaningful when that is BMPC_OK.  One synchronisation.
int bmpc_fold_shard_records(bmpc_ctx* ctx, int group, const void* d_records, size_t count, size_t stride,
                            uint8_t* out, uint32_t* flags_or_out, void* stream) {
    if (!ctx || !out || !flags_or_out || (!d_records && count)) return BMPC_ERR_INVALID;
    const size_t pb = bmpc_partial_bytes(group);
    if (pb == 0 || stride < pb + 4 || stride % 4 || count >= ((size_t)1 << 20)) return BMPC_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = pick_stream(ctx, stream);
    StreamScope ss(ctx, st);
    const size_t ob = group == BMPC_G1 ? 96 : 192;
    int rr = ws_reserve(ctx, ws_need(count * pb + 256, 1));
    if (rr) return rr;
    uint32_t* contig = ws_take<uint32_t>(ctx, count * pb / 4 + 4);
    if (!contig) return BMPC_ERR_INVALID;
    uint32_t* d_for = reinterpret_cast<uint32_t*>(ctx->d_stage + 32);
    uint8_t* d_bytes = ctx->d_stage + 64;
    CK(cudaMemsetAsync(d_for, 0, 4, st));
    if (count)
        LAUNCH(ctx, shard_records_unpack_kernel, (uint32_t)count, 64, 0, st, (const uint32_t*)d_records,
               (uint32_t)(stride / 4), (uint32_t)count, (uint32_t)(pb / 4), contig, d_for);
    int rcs = group == BMPC_G1 ? GroupOps<Fp>::sum_partials(ctx, contig, (uint32_t)count, d_bytes, st)
                               : GroupOps<Fp2>::sum_partials(ctx, contig, (uint32_t)count, d_bytes, st);
    if (rcs) return rcs;
    CK(cudaMemcpyAsync(ctx->h_stage, ctx->d_stage, 64 + ob, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *flags_or_out = *reinterpret_cast<const uint32_t*>(ctx->h_stage + 32);
    memcpy(out, ctx->h_stage + 64, ob);
    return BMPC_OK;
}

int bmpc_msm_flags_status(uint32_t flags_or) { return flags_to_status(flags_or); }

int bmpc_multiexp(bmpc_ctx* ctx, const bmpc_bases* bases, size_t base_offset, const uint64_t* scalars,
                  size_t n, const uint64_t* density_words, size_t density_len, uint8_t* out) {
    if (!ctx || !bases || !out || (!scalars && n)) return BMPC_ERR_INVALID;
    if (density_words && density_len != n) return BMPC_ERR_LENGTH_MISMATCH;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = ctx->own_stream;
    StreamScope ss(ctx, st);
    uint64_t* d_s = nullptr;
    uint64_t* d_d = nullptr;
    size_t dw = (n + 63) / 64;
    if (n) {
        size_t sbytes = align_up(n * 32, 256);
        int rr = io_reserve(ctx, sbytes + dw * 8 + 256);
        if (rr) return rr;
        d_s = reinterpret_cast<uint64_t*>(ctx->io);
        CK(cudaMemcpyAsync(d_s, scalars, n * 32, cudaMemcpyHostToDevice, st));
        if (density_words) {
            d_d = reinterpret_cast<uint64_t*>(ctx->io + sbytes);
            CK(cudaMemcpyAsync(d_d, density_words, dw * 8, cudaMemcpyHostToDevice, st));
        }
    }
    ctx->dense_hint = (density_words && n) ? popcount_bits(density_words, n) : 0;
    return multiexp_dev_locked(ctx, bases, base_offset, d_s, n, d_d, density_len, out, nullptr, st);
}

int bmpc_sum_partials(bmpc_ctx* ctx, int group, const void* d_partials, size_t count, uint8_t* out,
                      void* stream) {
    if (!ctx || !out || (!d_partials && count)) return BMPC_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = pick_stream(ctx, stream);
    StreamScope ss(ctx, st);
    size_t ob = group == BMPC_G1 ? 96 : 192;
    uint8_t* d_bytes = ctx->d_stage + 64;
    int rcs = group == BMPC_G1 ? GroupOps<Fp>::sum_partials(ctx, d_partials, (uint32_t)count, d_bytes, st)
                               : GroupOps<Fp2>::sum_partials(ctx, d_partials, (uint32_t)count, d_bytes, st);
    if (rcs) return rcs;
    CK(cudaMemcpyAsync(ctx->h_stage, d_bytes, ob, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    memcpy(out, ctx->h_stage, ob);
    return BMPC_OK;
}

// --------------------------------------------------------------------------- domain
static int domain_alloc(bmpc_ctx* ctx, size_t len, bmpc_domain** out) {
    // from_coeffs, domain.rs:47-60
    size_t m = 1;
    uint32_t exp = 0;
    while (m < len) {
        m *= 2;
        exp += 1;
        if (exp >= 32) return BMPC_ERR_DEGREE_TOO_LARGE;
    }
    bmpc_domain* d = new bmpc_domain();
    d->m = m;
    d->exp = exp;
    CK(cudaMalloc(&d->d, m * sizeof(Fr)));
    *out = d;
    return BMPC_OK;
}

int bmpc_domain_from_coeffs(bmpc_ctx* ctx, const uint64_t* coeffs, size_t len, bmpc_domain** out) {
    if (!ctx || !out || (!coeffs && len)) return BMPC_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = ctx->own_stream;
    StreamScope ss(ctx, st);
    bmpc_domain* d;
    int rc = domain_alloc(ctx, len, &d);
    if (rc) return rc;
    if (len) CK(cudaMemcpyAsync(d->d, coeffs, len * 32, cudaMemcpyHostToDevice, st));
    if (d->m > len) CK(cudaMemsetAsync(d->d + len, 0, (d->m - len) * 32, st));  // zero == Montgomery zero
    CK(cudaStreamSynchronize(st));
    *out = d;
    return BMPC_OK;
}

int bmpc_domain_from_coeffs_dev(bmpc_ctx* ctx, const uint64_t* d_coeffs, size_t len, bmpc_domain** out,
                                void* stream) {
    if (!ctx || !out || (!d_coeffs && len)) return BMPC_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = pick_stream(ctx, stream);
    StreamScope ss(ctx, st);
    bmpc_domain* d;
    int rc = domain_alloc(ctx, len, &d);
    if (rc) return rc;
    if (len) CK(cudaMemcpyAsync(d->d, d_coeffs, len * 32, cudaMemcpyDeviceToDevice, st));
    if (d->m > len) CK(cudaMemsetAsync(d->d + len, 0, (d->m - len) * 32, st));
    *out = d;
    return BMPC_OK;
}

size_t bmpc_domain_len(const bmpc_domain* d) { return d ? d->m : 0; }
uint32_t bmpc_domain_exp(const bmpc_domain* d) { return d ? d->exp : 0; }
uint64_t* bmpc_domain_dev_ptr(bmpc_domain* d) { return d ? reinterpret_cast<uint64_t*>(d->d) : nullptr; }

int bmpc_domain_into_coeffs(bmpc_ctx* ctx, const bmpc_domain* d, uint64_t* out) {
    if (!ctx || !d || !out) return BMPC_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    // the domain may still be in flight on a caller's stream: wait for the context's last call
    if (ctx->last_valid) CK(cudaEventSynchronize(ctx->last_done));
    CK(cudaStreamSynchronize(ctx->own_stream));
    CK(cudaMemcpy(out, d->d, d->m * 32, cudaMemcpyDeviceToHost));
    return BMPC_OK;
}

void bmpc_domain_free(bmpc_ctx* ctx, bmpc_domain* d) {
    if (!d) return;
    if (ctx) {
        DeviceGuard dg(ctx->device);
        cudaDeviceSynchronize();
        if (d->d) cudaFree(d->d);
    }
    delete d;
}

int bmpc_domain_transform(bmpc_ctx* ctx, bmpc_domain* d, int op, void* stream) {
    if (!ctx || !d) return BMPC_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = pick_stream(ctx, stream);
    StreamScope ss(ctx, st);
    return ntt_dev_locked(ctx, d->d, d->exp, op, st);
}

int bmpc_ntt_dev(bmpc_ctx* ctx, uint64_t* d_coeffs, uint32_t log_m, int op, void* stream) {
    if (!ctx || !d_coeffs) return BMPC_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = pick_stream(ctx, stream);
    StreamScope ss(ctx, st);
    return ntt_dev_locked(ctx, reinterpret_cast<Fr*>(d_coeffs), log_m, op, st);
}

int bmpc_ntt_batch_dev(bmpc_ctx* ctx, uint64_t* d_coeffs, uint32_t log_n, uint32_t batch, int inverse, void* stream) {
    if (!ctx || !d_coeffs) return BMPC_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = pick_stream(ctx, stream);
    StreamScope ss(ctx, st);
    return ntt_batch_dev_locked(ctx, reinterpret_cast<Fr*>(d_coeffs), log_n, batch, inverse != 0, st);
}

int bmpc_fr_swap01_dev(bmpc_ctx* ctx, const uint64_t* d_in, uint64_t* d_out, uint32_t d0, uint32_t d1, uint32_t d2,
                       void* stream) {
    if (!ctx || !d_in || !d_out || d_in == d_out) return BMPC_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = pick_stream(ctx, stream);
    StreamScope ss(ctx, st);
    return fr_swap01(ctx, reinterpret_cast<const Fr*>(d_in), reinterpret_cast<Fr*>(d_out), d0, d1, d2, st);
}

int bmpc_ntt_fourstep_twiddle_dev(bmpc_ctx* ctx, uint64_t* d, uint32_t rows, uint32_t cols, uint32_t row0,
                                  uint32_t log_m, int inverse, void* stream) {
    if (!ctx || !d) return BMPC_ERR_INVALID;
    if (log_m >= 32) return BMPC_ERR_DEGREE_TOO_LARGE;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = pick_stream(ctx, stream);
    StreamScope ss(ctx, st);
    return fr_fourstep_twiddle(ctx, reinterpret_cast<Fr*>(d), rows, cols, row0, log_m, inverse != 0, st);
}

int bmpc_fr_scale_pow_dev(bmpc_ctx* ctx, uint64_t* d, size_t n, uint32_t first, uint32_t log_m, int which,
                          void* stream) {
    if (!ctx || !d) return BMPC_ERR_INVALID;
    if (log_m >= 32) return BMPC_ERR_DEGREE_TOO_LARGE;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = pick_stream(ctx, stream);
    StreamScope ss(ctx, st);
    return fr_scale_pow(ctx, reinterpret_cast<Fr*>(d), n, first, log_m, which, st);
}

int bmpc_ntt(bmpc_ctx* ctx, uint64_t* coeffs, uint32_t log_m, int op) {
    if (!ctx || !coeffs) return BMPC_ERR_INVALID;
    if (log_m >= 32) return BMPC_ERR_DEGREE_TOO_LARGE;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = ctx->own_stream;
    StreamScope ss(ctx, st);
    size_t m = (size_t)1 << log_m;
    // staged through the context's grow-only device buffer: a prover calling fft() in a loop pays
    // no cudaMalloc / cudaFree per call
    int rr = io_reserve(ctx, m * sizeof(Fr));
    if (rr) return rr;
    Fr* d = reinterpret_cast<Fr*>(ctx->io);
    CK(cudaMemcpyAsync(d, coeffs, m * 32, cudaMemcpyHostToDevice, st));
    int rc = ntt_dev_locked(ctx, d, log_m, op, st);
    if (rc == BMPC_OK) {
        cudaError_t e = cudaMemcpyAsync(coeffs, d, m * 32, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = BMPC_ERR_CUDA; }
    }
    return rc;
}

int bmpc_domain_distribute_powers(bmpc_ctx* ctx, bmpc_domain* d, const uint64_t g[4], void* stream) {
    if (!ctx || !d || !g) return BMPC_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = pick_stream(ctx, stream);
    StreamScope ss(ctx, st);
    Fr* d_g = reinterpret_cast<Fr*>(ctx->d_stage + 1024);
    CK(cudaMemcpyAsync(d_g, g, 32, cudaMemcpyHostToDevice, st));
    int rcd = fr_distribute_powers(ctx, d->d, d->m, d_g, st);
    if (rcd) return rcd;
    CK(cudaStreamSynchronize(st));
    return BMPC_OK;
}

int bmpc_domain_z(bmpc_ctx* ctx, const bmpc_domain* d, const uint64_t tau[4], uint64_t out[4]) {
    if (!ctx || !d || !tau || !out) return BMPC_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = ctx->own_stream;
    StreamScope ss(ctx, st);
    Fr* d_t = reinterpret_cast<Fr*>(ctx->d_stage + 1024);
    CK(cudaMemcpyAsync(d_t, tau, 32, cudaMemcpyHostToDevice, st));
    int rcz = fr_eval_z(ctx, d_t, d->exp, d_t + 1, st);
    if (rcz) return rcz;
    CK(cudaMemcpyAsync(out, d_t + 1, 32, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return BMPC_OK;
}

int bmpc_domain_divide_by_z_on_coset(bmpc_ctx* ctx, bmpc_domain* d, void* stream) {
    if (!ctx || !d) return BMPC_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = pick_stream(ctx, stream);
    StreamScope ss(ctx, st);
    return fr_scale_zinv(ctx, d->d, d->m, d->exp, st);
}

int bmpc_domain_mul_assign(bmpc_ctx* ctx, bmpc_domain* d, const bmpc_domain* other, void* stream) {
    if (!ctx || !d || !other) return BMPC_ERR_INVALID;
    if (d->m != other->m) return BMPC_ERR_LENGTH_MISMATCH;  // assert_eq!, domain.rs:155
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = pick_stream(ctx, stream);
    StreamScope ss(ctx, st);
    return fr_pointwise(ctx, 0, d->d, other->d, d->m, st);
}

int bmpc_domain_sub_assign(bmpc_ctx* ctx, bmpc_domain* d, const bmpc_domain* other, void* stream) {
    if (!ctx || !d || !other) return BMPC_ERR_INVALID;
    if (d->m != other->m) return BMPC_ERR_LENGTH_MISMATCH;  // assert_eq!, domain.rs:174
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = pick_stream(ctx, stream);
    StreamScope ss(ctx, st);
    return fr_pointwise(ctx, 1, d->d, other->d, d->m, st);
}

// ---------------------------------------------------------------------- H polynomial
int bmpc_h_coefficients_dev(bmpc_ctx* ctx, uint64_t* d_a, uint64_t* d_b, uint64_t* d_c, uint32_t log_m,
                            void* stream) {
    if (!ctx || !d_a || !d_b || !d_c) return BMPC_ERR_INVALID;
    if (log_m >= 32) return BMPC_ERR_DEGREE_TOO_LARGE;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = pick_stream(ctx, stream);
    StreamScope ss(ctx, st);
    size_t m = (size_t)1 << log_m;
    int rc = ws_reserve(ctx, 2 * ws_need(m, sizeof(Fr)));
    if (rc) return rc;
    Fr* t1 = ws_take<Fr>(ctx, m);
    Fr* t2 = ws_take<Fr>(ctx, m);
    return h_coefficients_locked(ctx, (Fr*)d_a, (Fr*)d_b, (Fr*)d_c, log_m, t1, t2, st);
}

int bmpc_h_coset_evals_dev(bmpc_ctx* ctx, uint64_t* d_p, uint32_t log_m, const uint64_t* host_src, size_t host_len,
                           void* stream) {
    if (!ctx || !d_p) return BMPC_ERR_INVALID;
    if (log_m >= 32) return BMPC_ERR_DEGREE_TOO_LARGE;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = pick_stream(ctx, stream);
    StreamScope ss(ctx, st);
    size_t m = (size_t)1 << log_m;
    if (host_src) {      // from_coeffs (prover.rs:211-213): upload and pad with zeros
        if (host_len > m) return BMPC_ERR_DEGREE_TOO_LARGE;
        if (host_len) CK(cudaMemcpyAsync(d_p, host_src, host_len * 32, cudaMemcpyHostToDevice, st));
        if (m > host_len) CK(cudaMemsetAsync(reinterpret_cast<Fr*>(d_p) + host_len, 0, (m - host_len) * 32, st));
    }
    int rc = ws_reserve(ctx, 2 * ws_need(m, sizeof(Fr)));
    if (rc) return rc;
    Fr* t1 = ws_take<Fr>(ctx, m);
    Fr* t2 = ws_take<Fr>(ctx, m);
    return h_coset_evals_locked(ctx, (Fr*)d_p, log_m, t1, t2, st);
}

int bmpc_h_from_coset_evals_dev(bmpc_ctx* ctx, uint64_t* d_a, const uint64_t* d_b, const uint64_t* d_c,
                                uint32_t log_m, void* stream) {
    if (!ctx || !d_a || !d_b || !d_c) return BMPC_ERR_INVALID;
    if (log_m >= 32) return BMPC_ERR_DEGREE_TOO_LARGE;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = pick_stream(ctx, stream);
    StreamScope ss(ctx, st);
    size_t m = (size_t)1 << log_m;
    int rc = ws_reserve(ctx, 2 * ws_need(m, sizeof(Fr)));
    if (rc) return rc;
    Fr* t1 = ws_take<Fr>(ctx, m);
    Fr* t2 = ws_take<Fr>(ctx, m);
    return h_from_coset_evals_locked(ctx, (Fr*)d_a, (const Fr*)d_b, (const Fr*)d_c, log_m, t1, t2, st);
}

int bmpc_h_coefficients(bmpc_ctx* ctx, const uint64_t* a, const uint64_t* b, const uint64_t* c, size_t len,
                        uint64_t* out, size_t* out_len) {
    if (!ctx || !a || !b || !c || !out || !out_len) return BMPC_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = ctx->own_stream;
    StreamScope ss(ctx, st);
    size_t m = 1;
    uint32_t exp = 0;
    while (m < len) {
        m *= 2;
        exp++;
        if (exp >= 32) return BMPC_ERR_DEGREE_TOO_LARGE;
    }
    int rc = ws_reserve(ctx, 5 * ws_need(m, sizeof(Fr)));
    if (rc) return rc;
    Fr* d[3];
    const uint64_t* src[3] = {a, b, c};
    for (int k = 0; k < 3; k++) {
        d[k] = ws_take<Fr>(ctx, m);
        CK(cudaMemcpyAsync(d[k], src[k], len * 32, cudaMemcpyHostToDevice, st));
        if (m > len) CK(cudaMemsetAsync(d[k] + len, 0, (m - len) * 32, st));
    }
    Fr* t1 = ws_take<Fr>(ctx, m);
    Fr* t2 = ws_take<Fr>(ctx, m);
    rc = h_coefficients_locked(ctx, d[0], d[1], d[2], exp, t1, t2, st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, d[0], (m - 1) * 32, cudaMemcpyDeviceToHost, st));  // prover.rs:227-229
    CK(cudaStreamSynchronize(st));
    *out_len = m - 1;
    return BMPC_OK;
}

int bmpc_fr_to_canonical_dev(bmpc_ctx* ctx, uint64_t* d_vals, size_t n, void* stream) {
    if (!ctx || (!d_vals && n)) return BMPC_ERR_INVALID;
    if (!n) return BMPC_OK;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = pick_stream(ctx, stream);
    StreamScope ss(ctx, st);
    return fr_pointwise(ctx, 2, (Fr*)d_vals, nullptr, n, st);
}

// ----------------------------------------------------------------------- create_proof
}  // extern "C"
namespace bmpc {
// Shared body of bmpc_create_proof (slices == NULL, partials_out == NULL: the eight multiexps over
// their whole exponent vectors, then the tail, 192-byte proof) and of one device's / rank's share of
// a sharded proof (partials_out != NULL: `S` is the assignment the positions of `slices` index into,
// `P` the device's slices of the query vectors; the XYZZ partial sums go to partials_out, the
// per-multiexp raw flag words to flags_out, no tail).
int create_proof_common(bmpc_ctx* ctx, const bmpc_params* P, const bmpc_assignment* S, const uint64_t r[4],
                        const uint64_t s[4], const ProofSlices* slices, uint8_t* proof_out,
                        uint8_t* partials_out, uint32_t* flags_out) {
    if (!P->h || !P->l || !P->a || !P->b_g1 || !P->b_g2) return BMPC_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = ctx->own_stream;
    StreamScope ss(ctx, st);
    const bool partial_mode = partials_out != nullptr;
    const size_t nc = S->num_constraints, ni = S->num_inputs, na = S->num_aux;
    size_t m = 1;
    uint32_t exp = 0;
    while (m < nc) {
        m *= 2;
        exp++;
        if (exp >= 32) return BMPC_ERR_DEGREE_TOO_LARGE;  // from_coeffs, prover.rs:211
    }
    // the eight multiexps (prover.rs:233,252-307) over their whole exponent vectors unless sliced
    ProofSlices whole;
    if (!slices) {
        size_t b_in_total = 0;
        for (size_t i = 0; i < ni; i++) b_in_total += (S->b_input_density[i / 64] >> (i % 64)) & 1;
        const size_t his[8] = {ni, na, ni, na, ni, na, m - 1, na};
        const size_t offs[8] = {0, ni, 0, b_in_total, 0, b_in_total, 0, 0};
        const uint64_t* ds[8] = {nullptr, S->a_aux_density, S->b_input_density, S->b_aux_density,
                                 S->b_input_density, S->b_aux_density, nullptr, nullptr};
        for (int j = 0; j < 8; j++) {
            whole.lo[j] = 0; whole.hi[j] = his[j]; whole.base_offset[j] = offs[j]; whole.n_total[j] = 0;
            whole.dens[j] = ds[j];
        }
        slices = &whole;
    }
    const ProofSlices& SL = *slices;
    {
        const size_t lim[8] = {ni, na, ni, na, ni, na, m - 1, na};
        for (int j = 0; j < 8; j++)
            if (SL.lo[j] > SL.hi[j] || SL.hi[j] > lim[j]) return BMPC_ERR_INVALID;
    }
    // the ranges of the input / aux assignment this call touches
    auto span = [&](const int* js, int cnt, size_t* lo, size_t* hi) {
        *lo = *hi = 0;
        bool any = false;
        for (int k = 0; k < cnt; k++) {
            const int j = js[k];
            if (SL.hi[j] == SL.lo[j]) continue;
            if (!any) { *lo = SL.lo[j]; *hi = SL.hi[j]; any = true; }
            else { if (SL.lo[j] < *lo) *lo = SL.lo[j]; if (SL.hi[j] > *hi) *hi = SL.hi[j]; }
        }
    };
    const int in_jobs[3] = {0, 2, 4}, aux_jobs[4] = {1, 3, 5, 7};
    size_t in_lo, in_hi, aux_lo, aux_hi;
    span(in_jobs, 3, &in_lo, &in_hi);
    span(aux_jobs, 4, &aux_lo, &aux_hi);
    const size_t n_in = in_hi - in_lo, n_aux = aux_hi - aux_lo;
    // Device buffers for this proof come out of the grow-only staging area (no cudaMalloc/cudaFree
    // per proof).  The big a, b, c upload runs on a second stream and overlaps the seven
    // multiexps that only need the assignments; the H pipeline waits for it.
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += align_up(bytes ? bytes : 1, 256); return o; };
    const size_t o_a = carve(m * 32), o_b = carve(m * 32), o_c = carve(m * 32);
    const size_t o_in = carve(n_in * 32), o_aux = carve(n_aux * 32);
    // density words per multiexp; jobs that read the same words over the same positions share a buffer
    size_t o_dens[8];
    int dens_src[8];
    for (int j = 0; j < 8; j++) {
        o_dens[j] = 0;
        dens_src[j] = -1;
        if (!SL.dens[j] || SL.hi[j] == SL.lo[j]) continue;
        dens_src[j] = j;
        for (int k = 0; k < j; k++)
            if (dens_src[k] == k && SL.dens[k] == SL.dens[j] && SL.lo[k] == SL.lo[j] && SL.hi[k] == SL.hi[j]) dens_src[j] = k;
        if (dens_src[j] == j) o_dens[j] = carve((SL.hi[j] - SL.lo[j] + 63) / 64 * 8);
        else o_dens[j] = o_dens[dens_src[j]];
    }
    const size_t o_misc = carve(8192);
    {
        int rr = io_reserve(ctx, off);
        if (rr) return rr;
    }
    Fr* d_a = reinterpret_cast<Fr*>(ctx->io + o_a);
    Fr* d_b = reinterpret_cast<Fr*>(ctx->io + o_b);
    Fr* d_c = reinterpret_cast<Fr*>(ctx->io + o_c);
    Fr* d_in = reinterpret_cast<Fr*>(ctx->io + o_in);
    Fr* d_aux = reinterpret_cast<Fr*>(ctx->io + o_aux);
    uint8_t* d_misc = reinterpret_cast<uint8_t*>(ctx->io + o_misc);  // partials + vk + r,s + proof
    if (!ctx->copy_stream) {
        CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&ctx->copy_done, cudaEventDisableTiming));
    }
    cudaStream_t cs = ctx->copy_stream;
    auto cleanup = [&]() {
        cudaStreamSynchronize(cs);
        for (auto& sl : ctx->slots)
            if (sl.stream) cudaStreamSynchronize(sl.stream);
        cudaStreamSynchronize(st);
    };
#define CKP(call)                                                            \
    do {                                                                     \
        cudaError_t e_ = (call);                                             \
        if (e_ != cudaSuccess) {                                             \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);   \
            cleanup();                                                       \
            return BMPC_ERR_CUDA;                                            \
        }                                                                    \
    } while (0)
#define RCP(expr)                  \
    do {                           \
        int rc_ = (expr);          \
        if (rc_ != BMPC_OK) {      \
            cleanup();             \
            return rc_;            \
        }                          \
    } while (0)
    // small inputs first, on the compute stream
    if (n_in) CKP(cudaMemcpyAsync(d_in, S->input_assignment + 4 * in_lo, n_in * 32, cudaMemcpyHostToDevice, st));
    if (n_aux) CKP(cudaMemcpyAsync(d_aux, S->aux_assignment + 4 * aux_lo, n_aux * 32, cudaMemcpyHostToDevice, st));
    for (int j = 0; j < 8; j++)
        if (dens_src[j] == j)
            CKP(cudaMemcpyAsync(ctx->io + o_dens[j], SL.dens[j], (SL.hi[j] - SL.lo[j] + 63) / 64 * 8,
                                cudaMemcpyHostToDevice, st));
    // evaluation vectors on the copy stream
    const Fr* srcs[3] = {(const Fr*)S->a, (const Fr*)S->b, (const Fr*)S->c};
    Fr* dsts[3] = {d_a, d_b, d_c};
    const bool want_h = SL.hi[6] > SL.lo[6];
    if (want_h) {
        for (int k = 0; k < 3; k++) {
            if (nc) CKP(cudaMemcpyAsync(dsts[k], srcs[k], nc * 32, cudaMemcpyHostToDevice, cs));
            if (m > nc) CKP(cudaMemsetAsync(dsts[k] + nc, 0, (m - nc) * 32, cs));
        }
    }
    CKP(cudaEventRecord(ctx->copy_done, cs));

    // layout of d_misc
    G1XYZZ* part_g1 = reinterpret_cast<G1XYZZ*>(d_misc);              // 6 x 192
    G2XYZZ* part_g2 = reinterpret_cast<G2XYZZ*>(d_misc + 1536);       // 2 x 384
    G1Affine* vk_g1 = reinterpret_cast<G1Affine*>(d_misc + 2560);     // 3 x 96
    G2Affine* vk_g2 = reinterpret_cast<G2Affine*>(d_misc + 3072);     // 2 x 192
    Fr* d_rs = reinterpret_cast<Fr*>(d_misc + 3584);                  // 2 x 32
    uint8_t* d_proof = d_misc + 3840;                                 // 192
    uint8_t* d_vkraw = d_misc + 4096;                                 // 672
    uint8_t vkraw[672];
    memcpy(vkraw, P->alpha_g1, 96); memcpy(vkraw + 96, P->beta_g1, 96); memcpy(vkraw + 192, P->delta_g1, 96);
    memcpy(vkraw + 288, P->beta_g2, 192); memcpy(vkraw + 480, P->delta_g2, 192);
    if (!partial_mode) {
        CKP(cudaMemcpyAsync(d_vkraw, vkraw, 672, cudaMemcpyHostToDevice, st));
        CKP(cudaMemcpyAsync(d_rs, r, 32, cudaMemcpyHostToDevice, st));
        CKP(cudaMemcpyAsync(d_rs + 1, s, 32, cudaMemcpyHostToDevice, st));
        RCP(GroupOps<Fp>::decode(ctx, d_vkraw, 96, 3, vk_g1, st));
        RCP(GroupOps<Fp2>::decode(ctx, d_vkraw + 288, 192, 2, vk_g2, st));
    }

    // to_le_bits of the assignments (prover.rs:237-250)
    RCP(fr_pointwise(ctx, 2, d_in, nullptr, n_in, st));
    RCP(fr_pointwise(ctx, 2, d_aux, nullptr, n_aux, st));

    struct Job { const bmpc_bases* bases; const uint64_t* sc; void* out; };
    auto in_sc = [&](int j) { return (const uint64_t*)(d_in + (SL.hi[j] > SL.lo[j] ? SL.lo[j] - in_lo : 0)); };
    auto aux_sc = [&](int j) { return (const uint64_t*)(d_aux + (SL.hi[j] > SL.lo[j] ? SL.lo[j] - aux_lo : 0)); };
    Job jobs[8] = {
        {P->a, in_sc(0), part_g1 + 0},        // a_inputs   :264-269
        {P->a, aux_sc(1), part_g1 + 1},       // a_aux      :270-275
        {P->b_g1, in_sc(2), part_g1 + 2},     // b_g1_inputs:285-290
        {P->b_g1, aux_sc(3), part_g1 + 3},    // b_g1_aux   :291-296
        {P->b_g2, in_sc(4), part_g2 + 0},     // b_g2_inputs:301-306
        {P->b_g2, aux_sc(5), part_g2 + 1},    // b_g2_aux   :307
        {P->h, (const uint64_t*)(d_a + SL.lo[6]), part_g1 + 4},   // h :233
        {P->l, aux_sc(7), part_g1 + 5},       // l          :252-257
    };
    // All eight are enqueued before anything is awaited, as three chains on three streams, each
    // with its own scratch arena and staging words (slot 0 = the context's own):
    //   slot 1: the G2 multiexps            slot 2: H pipeline + H, then the B-G1 multiexps
    //   slot 0: the A multiexps and L
    // so that one chain's multiplier-pipe bubbles (sort: atomics-bound; bucket reduction and the
    // transforms' tails: latency-bound) are filled by the others' accumulate kernels.  Within a
    // slot the multiexps follow each other in stream order and reuse the arena.  Flags / results
    // of job j land at h_stage + 256 j.  (One chain, one wait per multiexp: 155 ms at 2^22; G2 on a
    // second stream: 123 ms.)
    int statuses[8];
    MsmPending pend[8];
    const int order[8] = {4, 5, 6, 0, 1, 7, 2, 3};
    const int slot_of3[8] = {0, 0, 2, 2, 1, 1, 2, 0};
    // Eight slots: every multiexp on its own stream and arena, so that one multiexp's bucket reduction
    // (a serial chain of ~50 additions on a few warps) runs under the next one's accumulation: 2^22
    // 0.1213 -> 0.1195 s, one rank's share of 8: 31.9 -> 30.5 ms.  Eight arenas cost memory (~5 GB each
    // at 2^22), so the automatic choice takes them up to 2^23 constraints only.
    const int slot_of8[8] = {0, 1, 2, 3, 4, 5, 6, 7};
    const bool eight = ctx->tune.proof_slots == 8 || (ctx->tune.proof_slots == 0 && exp <= 23);
    const int* slot_of = eight ? slot_of8 : slot_of3;
    if (!ctx->slots[0].stream) {
        for (auto& sl : ctx->slots) {
            CKP(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
            CKP(cudaMalloc(&sl.d_stage, 4096));
        }
        CKP(cudaEventCreateWithFlags(&ctx->inputs_ready, cudaEventDisableTiming));
    }
    CKP(cudaEventRecord(ctx->inputs_ready, st));
    for (auto& sl : ctx->slots) CKP(cudaStreamWaitEvent(sl.stream, ctx->inputs_ready, 0));
    auto swap_slot = [&](int k) {    // slot 0 <-> slot k (k = 1, 2): scratch arena and device staging
        bmpc_ctx::Slot& sl = ctx->slots[k - 1];
        std::swap(ctx->ws, sl.ws);
        std::swap(ctx->ws_size, sl.ws_size);
        std::swap(ctx->d_stage, sl.d_stage);
        ctx->ws_used = 0;
    };
    for (int k = 0; k < 8; k++) {
        const int j = order[k], sk = slot_of[j];
        cudaStream_t js = sk ? ctx->slots[sk - 1].stream : st;
        if (sk) swap_slot(sk);
        int rc = BMPC_OK;
        if (j == 6 && want_h) {
            // H polynomial (prover.rs:210-231)
            cudaError_t e_ = cudaStreamWaitEvent(js, ctx->copy_done, 0);
            if (e_ != cudaSuccess) rc = BMPC_ERR_CUDA;
            if (!rc) rc = ws_reserve(ctx, 2 * ws_need(m, sizeof(Fr)));
            if (!rc) {
                Fr* t1 = ws_take<Fr>(ctx, m);
                Fr* t2 = ws_take<Fr>(ctx, m);
                rc = h_coefficients_locked(ctx, d_a, d_b, d_c, exp, t1, t2, js);
            }
        }
        if (!rc) {
            const size_t nj = SL.hi[j] - SL.lo[j];
            const uint64_t* dj = dens_src[j] >= 0 ? reinterpret_cast<const uint64_t*>(ctx->io + o_dens[j]) : nullptr;
            ctx->dense_hint = dj ? popcount_bits(SL.dens[j], nj) : 0;
            rc = multiexp_enqueue(ctx, jobs[j].bases, SL.base_offset[j], jobs[j].sc, nj, dj, nj, nullptr, jobs[j].out,
                                  js, ctx->h_stage + 256 * j, &pend[j], SL.n_total[j]);
        }
        if (sk) swap_slot(sk);
        if (rc != BMPC_OK) {      // argument / launch errors; the multiexp statuses come from the flags
            cleanup();
            return rc;
        }
    }
    for (auto& sl : ctx->slots) CKP(cudaStreamSynchronize(sl.stream));
    CKP(cudaStreamSynchronize(st));
    uint32_t flags[8];
    for (int j = 0; j < 8; j++) statuses[j] = multiexp_collect(pend[j], &flags[j]);
    if (partial_mode) {   // one share: hand back the partial sums and raw flags, the caller gathers them
        CKP(cudaMemcpyAsync(ctx->h_stage, part_g1, 2304, cudaMemcpyDeviceToHost, st));
        CKP(cudaStreamSynchronize(st));
        memcpy(partials_out, ctx->h_stage, 6 * sizeof(G1XYZZ));
        memcpy(partials_out + 6 * sizeof(G1XYZZ), ctx->h_stage + 1536, 2 * sizeof(G2XYZZ));
        for (int j = 0; j < 8; j++) flags_out[j] = flags[j];
        cleanup();
        return BMPC_OK;
    }
    // subversion check delta != identity comes before the first wait() (prover.rs:309-313)
    if ((P->delta_g1[0] & 0x40) || (P->delta_g2[0] & 0x40)) {
        cleanup();
        return BMPC_ERR_UNEXPECTED_IDENTITY;
    }
    for (int j = 0; j < 8; j++)
        if (statuses[j] != BMPC_OK) {
            cleanup();
            return statuses[j];
        }
    ProveTailArgs T;
    T.a_inputs = part_g1 + 0; T.a_aux = part_g1 + 1; T.b1_inputs = part_g1 + 2; T.b1_aux = part_g1 + 3;
    T.b2_inputs = part_g2 + 0; T.b2_aux = part_g2 + 1; T.h = part_g1 + 4; T.l = part_g1 + 5;
    T.vk_g1 = vk_g1; T.vk_g2 = vk_g2; T.rs = d_rs; T.proof = d_proof;
    RCP(prove_tail_launch(ctx, T, st));
    CKP(cudaMemcpyAsync(ctx->h_stage, d_proof, 192, cudaMemcpyDeviceToHost, st));
    CKP(cudaStreamSynchronize(st));
    memcpy(proof_out, ctx->h_stage, 192);
    cleanup();
    return BMPC_OK;
#undef CKP
#undef RCP
}
}  // namespace bmpc

extern "C" {
int bmpc_create_proof(bmpc_ctx* ctx, const bmpc_params* P, const bmpc_assignment* S, const uint64_t r[4],
                      const uint64_t s[4], uint8_t proof_out[192]) {
    if (!ctx || !P || !S || !r || !s || !proof_out) return BMPC_ERR_INVALID;
    return create_proof_common(ctx, P, S, r, s, nullptr, proof_out, nullptr, nullptr);
}

int bmpc_create_proof_partials(bmpc_ctx* ctx, const bmpc_params* P, const bmpc_assignment* S,
                               const bmpc_proof_shard* shard, uint8_t partials_out[BMPC_PROOF_PARTIAL_BYTES],
                               uint32_t flags_out[8]) {
    if (!ctx || !P || !S || !shard || !partials_out || !flags_out) return BMPC_ERR_INVALID;
    // `S` is the rank's slice of the assignment: every multiexp covers all of it
    ProofSlices sl;
    const size_t ni = S->num_inputs, na = S->num_aux;
    const size_t his[8] = {ni, na, ni, na, ni, na, shard->h_hi, na};
    const uint64_t* ds[8] = {nullptr, S->a_aux_density, S->b_input_density, S->b_aux_density,
                             S->b_input_density, S->b_aux_density, nullptr, nullptr};
    for (int j = 0; j < 8; j++) {
        sl.lo[j] = j == 6 ? shard->h_lo : 0;
        sl.hi[j] = his[j];
        sl.base_offset[j] = shard->base_offset[j];
        sl.n_total[j] = shard->n_total[j];
        sl.dens[j] = ds[j];
    }
    return create_proof_common(ctx, P, S, nullptr, nullptr, &sl, nullptr, partials_out, flags_out);
}

int bmpc_create_proof_finish(bmpc_ctx* ctx, const bmpc_params* P, const uint8_t* partials_all, size_t world,
                             const uint64_t r[4], const uint64_t s[4], uint8_t proof_out[192]) {
    if (!ctx || !P || !partials_all || !world || world > 64 || !r || !s || !proof_out) return BMPC_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = ctx->own_stream;
    StreamScope ss(ctx, st);
    // subversion check first, as in the single-GPU path (prover.rs:309-313)
    if ((P->delta_g1[0] & 0x40) || (P->delta_g2[0] & 0x40)) return BMPC_ERR_UNEXPECTED_IDENTITY;
    const size_t need = align_up(world * BMPC_PROOF_PARTIAL_BYTES, 256) + 8192;
    int rr = io_reserve(ctx, need);
    if (rr) return rr;
    uint8_t* d_all = reinterpret_cast<uint8_t*>(ctx->io);
    uint8_t* d_misc = d_all + align_up(world * BMPC_PROOF_PARTIAL_BYTES, 256);
    G1XYZZ* part_g1 = reinterpret_cast<G1XYZZ*>(d_misc);
    G2XYZZ* part_g2 = reinterpret_cast<G2XYZZ*>(d_misc + 1536);
    G1Affine* vk_g1 = reinterpret_cast<G1Affine*>(d_misc + 2560);
    G2Affine* vk_g2 = reinterpret_cast<G2Affine*>(d_misc + 3072);
    Fr* d_rs = reinterpret_cast<Fr*>(d_misc + 3584);
    uint8_t* d_proof = d_misc + 3840;
    uint8_t* d_vkraw = d_misc + 4096;
    uint8_t vkraw[672];
    memcpy(vkraw, P->alpha_g1, 96); memcpy(vkraw + 96, P->beta_g1, 96); memcpy(vkraw + 192, P->delta_g1, 96);
    memcpy(vkraw + 288, P->beta_g2, 192); memcpy(vkraw + 480, P->delta_g2, 192);
    CK(cudaMemcpyAsync(d_all, partials_all, world * BMPC_PROOF_PARTIAL_BYTES, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_vkraw, vkraw, 672, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_rs, r, 32, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_rs + 1, s, 32, cudaMemcpyHostToDevice, st));
    int rc = GroupOps<Fp>::decode(ctx, d_vkraw, 96, 3, vk_g1, st);
    if (rc) return rc;
    rc = GroupOps<Fp2>::decode(ctx, d_vkraw + 288, 192, 2, vk_g2, st);
    if (rc) return rc;
    rc = prove_fold_partials_launch(ctx, d_all, (uint32_t)world, part_g1, part_g2, st);
    if (rc) return rc;
    ProveTailArgs T;
    T.a_inputs = part_g1 + 0; T.a_aux = part_g1 + 1; T.b1_inputs = part_g1 + 2; T.b1_aux = part_g1 + 3;
    T.b2_inputs = part_g2 + 0; T.b2_aux = part_g2 + 1; T.h = part_g1 + 4; T.l = part_g1 + 5;
    T.vk_g1 = vk_g1; T.vk_g2 = vk_g2; T.rs = d_rs; T.proof = d_proof;
    rc = prove_tail_launch(ctx, T, st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->h_stage, d_proof, 192, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    memcpy(proof_out, ctx->h_stage, 192);
    return BMPC_OK;
}

// ------------------------------------------------------------- Parameters wire format
namespace {

// Decode + validate `count` points starting at data[pos]; fewer may be available (short blob).
// Returns BMPC_OK, BMPC_ERR_INVALID_DATA (ctx->err set) or BMPC_ERR_UNEXPECTED_EOF, in the order
// the reference's sequential reader would hit them.
int read_section(bmpc_ctx* ctx, int group, const uint8_t* data, size_t len, size_t* pos, size_t count,
                 int checked, int reject_identity, bmpc_bases** out, cudaStream_t st) {
    const size_t pb = group == BMPC_G1 ? 96 : 192;
    size_t avail = (len - *pos) / pb;
    size_t n = count < avail ? count : avail;
    BasesGuard bg(ctx);
    bmpc_bases* b = bg.b = new bmpc_bases();
    b->group = group;
    b->n = n;
    CK(cudaMalloc(&b->d_points, (n ? n : 1) * pb));
    uint32_t* d_err = reinterpret_cast<uint32_t*>(ctx->d_stage + 1536);
    uint32_t h_err = 0xffffffffu;
    DevBuf raw;
    if (n) {
        CK(cudaMalloc(&raw.p, n * pb));
        uint8_t* d_raw = raw.as<uint8_t>();
        CK(cudaMemcpyAsync(d_raw, data + *pos, n * pb, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_err, &h_err, 4, cudaMemcpyHostToDevice, st));
        int rc = group == BMPC_G1
                     ? GroupOps<Fp>::validate_decode(ctx, d_raw, pb, n, checked, reject_identity, b->d_points, d_err, st)
                     : GroupOps<Fp2>::validate_decode(ctx, d_raw, pb, n, checked, reject_identity, b->d_points, d_err, st);
        if (rc) return rc;
        CK(cudaMemcpyAsync(&h_err, d_err, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    *pos += n * pb;
    if (h_err != 0xffffffffu) {
        ctx->err = (h_err & 3u) == 2u ? "point at infinity" : (group == BMPC_G1 ? "invalid G1" : "invalid G2");
        return BMPC_ERR_INVALID_DATA;
    }
    if (n < count) return BMPC_ERR_UNEXPECTED_EOF;
    size_t nw = (n + 31) / 32 + 1;
    CK(cudaMalloc(&b->d_inf, nw * 4));
    CK(cudaMemsetAsync(b->d_inf, 0, nw * 4, st));
    int rc = group == BMPC_G1 ? GroupOps<Fp>::inf_bitmap(ctx, b->d_points, n, b->d_inf, st)
                              : GroupOps<Fp2>::inf_bitmap(ctx, b->d_points, n, b->d_inf, st);
    if (rc) return rc;
    CK(cudaStreamSynchronize(st));
    *out = bg.release();
    return BMPC_OK;
}

int read_u32_be(const uint8_t* data, size_t len, size_t* pos, size_t* out) {
    if (*pos + 4 > len) return BMPC_ERR_UNEXPECTED_EOF;
    const uint8_t* p = data + *pos;
    *out = ((size_t)p[0] << 24) | ((size_t)p[1] << 16) | ((size_t)p[2] << 8) | p[3];
    *pos += 4;
    return BMPC_OK;
}

}  // namespace

void bmpc_params_free(bmpc_ctx* ctx, bmpc_parameters* p) {
    if (!p) return;
    const bmpc_bases** hs[5] = {&p->p.h, &p->p.l, &p->p.a, &p->p.b_g1, &p->p.b_g2};
    for (auto h : hs) {
        if (*h) bmpc_bases_free(ctx, const_cast<bmpc_bases*>(*h));
        *h = nullptr;
    }
    if (p->ic) bmpc_bases_free(ctx, p->ic);
    p->ic = nullptr;
}

int bmpc_params_read(bmpc_ctx* ctx, const uint8_t* data, size_t len, int checked, bmpc_parameters* out) {
    if (!ctx || !data || !out) return BMPC_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = ctx->own_stream;
    StreamScope ss(ctx, st);
    memset(out, 0, sizeof(*out));
    size_t pos = 0;
    // VerifyingKey::read (mod.rs:161-221): six points, always checked, identity allowed
    struct { int group; uint8_t* dst; } vk[6] = {
        {BMPC_G1, out->p.alpha_g1}, {BMPC_G1, out->p.beta_g1}, {BMPC_G2, out->p.beta_g2},
        {BMPC_G2, out->gamma_g2},   {BMPC_G1, out->p.delta_g1}, {BMPC_G2, out->p.delta_g2}};
    int rc = BMPC_OK;
    for (int k = 0; k < 6 && rc == BMPC_OK; k++) {
        size_t pb = vk[k].group == BMPC_G1 ? 96 : 192, before = pos;
        bmpc_bases* tmp = nullptr;
        rc = read_section(ctx, vk[k].group, data, len, &pos, 1, 1, 0, &tmp, st);
        if (rc == BMPC_OK) {
            memcpy(vk[k].dst, data + before, pb);
            bmpc_bases_free(ctx, tmp);
        }
    }
    size_t n = 0;
    if (rc == BMPC_OK) rc = read_u32_be(data, len, &pos, &n);
    if (rc == BMPC_OK) rc = read_section(ctx, BMPC_G1, data, len, &pos, n, 1, 1, &out->ic, st);   // ic: checked, no identity
    // Parameters::read (mod.rs:292-400): h, l, a, b_g1 (G1), b_g2 (G2); identity rejected
    struct { int group; const bmpc_bases** dst; } q[5] = {
        {BMPC_G1, &out->p.h}, {BMPC_G1, &out->p.l}, {BMPC_G1, &out->p.a}, {BMPC_G1, &out->p.b_g1}, {BMPC_G2, &out->p.b_g2}};
    for (int k = 0; k < 5 && rc == BMPC_OK; k++) {
        rc = read_u32_be(data, len, &pos, &n);
        bmpc_bases* b = nullptr;
        if (rc == BMPC_OK) rc = read_section(ctx, q[k].group, data, len, &pos, n, checked, 1, &b, st);
        if (rc == BMPC_OK) *q[k].dst = b;
    }
    if (rc != BMPC_OK) {
        std::string keep = ctx->err;
        bmpc_params_free(ctx, out);
        ctx->err = keep;
    }
    return rc;
}

int bmpc_params_write(bmpc_ctx* ctx, const bmpc_parameters* in, uint8_t* out, size_t cap, size_t* written) {
    if (!ctx || !in || !written || !in->ic || !in->p.h || !in->p.l || !in->p.a || !in->p.b_g1 || !in->p.b_g2)
        return BMPC_ERR_INVALID;
    const bmpc_bases* q[6] = {in->ic, in->p.h, in->p.l, in->p.a, in->p.b_g1, in->p.b_g2};
    size_t need = 864;
    for (auto b : q) need += 4 + b->n * (b->group == BMPC_G1 ? 96 : 192);
    *written = need;
    if (!out || cap < need) return BMPC_ERR_INVALID;
    uint8_t* p = out;
    memcpy(p, in->p.alpha_g1, 96); p += 96;
    memcpy(p, in->p.beta_g1, 96); p += 96;
    memcpy(p, in->p.beta_g2, 192); p += 192;
    memcpy(p, in->gamma_g2, 192); p += 192;
    memcpy(p, in->p.delta_g1, 96); p += 96;
    memcpy(p, in->p.delta_g2, 192); p += 192;
    for (auto b : q) {
        p[0] = (uint8_t)(b->n >> 24); p[1] = (uint8_t)(b->n >> 16); p[2] = (uint8_t)(b->n >> 8); p[3] = (uint8_t)b->n;
        p += 4;
        int rc = bmpc_bases_read(ctx, b, 0, b->n, p);
        if (rc) return rc;
        p += b->n * (b->group == BMPC_G1 ? 96 : 192);
    }
    return BMPC_OK;
}

// -------------------------------------------------------------- batch scalar multiply
int bmpc_batch_scalar_mul(bmpc_ctx* ctx, const bmpc_bases* in, const uint64_t* scalars, int per_element,
                          bmpc_bases** out) {
    if (!ctx || !in || !scalars || !out) return BMPC_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = ctx->own_stream;
    StreamScope ss(ctx, st);
    size_t n = in->n;
    size_t pb = in->group == BMPC_G1 ? 96 : 192;
    size_t ns = per_element ? n : 1;
    DevBuf sbuf;
    BasesGuard bg(ctx);
    CK(cudaMalloc(&sbuf.p, (ns ? ns : 1) * 32));
    uint32_t* d_s = sbuf.as<uint32_t>();
    if (ns) CK(cudaMemcpyAsync(d_s, scalars, ns * 32, cudaMemcpyHostToDevice, st));
    bmpc_bases* b = bg.b = new bmpc_bases();
    b->group = in->group;
    b->n = n;
    CK(cudaMalloc(&b->d_points, (n ? n : 1) * pb));
    int rcb = in->group == BMPC_G1 ? GroupOps<Fp>::batch_mul(ctx, in->d_points, d_s, per_element, n, b->d_points, st)
                                   : GroupOps<Fp2>::batch_mul(ctx, in->d_points, d_s, per_element, n, b->d_points, st);
    if (rcb) return rcb;
    int rc = register_points(ctx, b, st);
    if (rc) return rc;
    CK(cudaStreamSynchronize(st));
    *out = bg.release();
    return BMPC_OK;
}

// -------------------------------------------------------------- list x sparse matrix (mpc.rs:416-457)
int bmpc_list_mul_matrix(bmpc_ctx* ctx, const bmpc_bases* list, const uint64_t* row_ptr, const uint32_t* cols,
                         const uint64_t* coeffs, size_t n_rows, bmpc_bases** out) {
    if (!ctx || !list || !row_ptr || !out) return BMPC_ERR_INVALID;
    const size_t n = list->n;
    size_t live = 0;                                     // rows before the first empty one (:432-434)
    while (live < n_rows && row_ptr[live + 1] > row_ptr[live]) live++;
    // result[i] indexes a vector of list.len() elements (:428-429, :445) for the rows the loop
    // PROCESSES: the reference panics only if a live row index reaches list.len() -- a matrix taller
    // than the list whose first empty row comes earlier is fine -- or on a live column >= list.len()
    if (live > n) return BMPC_ERR_LENGTH_MISMATCH;
    const size_t j0 = row_ptr[0], nnz = row_ptr[live] - j0;
    if (nnz && (!cols || !coeffs)) return BMPC_ERR_INVALID;
    if (nnz >= ((size_t)1 << 32)) return BMPC_ERR_INVALID;
    for (size_t j = 0; j < nnz; j++)
        if (cols[j0 + j] >= n) return BMPC_ERR_LENGTH_MISMATCH;
    std::vector<uint32_t> rp(live + 1);
    for (size_t i = 0; i <= live; i++) rp[i] = (uint32_t)(row_ptr[i] - j0);
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = ctx->own_stream;
    StreamScope ss(ctx, st);
    const size_t pb = list->group == BMPC_G1 ? 96 : 192;
    DevBuf rp_buf, col_buf, cf_buf;       // freed on every exit path
    BasesGuard bg(ctx);
    CK(cudaMalloc(&rp_buf.p, (live + 1) * 4));
    CK(cudaMalloc(&col_buf.p, (nnz ? nnz : 1) * 4));
    CK(cudaMalloc(&cf_buf.p, (nnz ? nnz : 1) * 32));
    uint32_t *d_rp = rp_buf.as<uint32_t>(), *d_col = col_buf.as<uint32_t>(), *d_cf = cf_buf.as<uint32_t>();
    CK(cudaMemcpyAsync(d_rp, rp.data(), (live + 1) * 4, cudaMemcpyHostToDevice, st));
    if (nnz) {
        CK(cudaMemcpyAsync(d_col, cols + j0, nnz * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_cf, coeffs + j0 * 4, nnz * 32, cudaMemcpyHostToDevice, st));
    }
    bmpc_bases* b = bg.b = new bmpc_bases();
    b->group = list->group;
    b->n = n;
    CK(cudaMalloc(&b->d_points, (n ? n : 1) * pb));
    int rcb = list->group == BMPC_G1
                  ? GroupOps<Fp>::list_mul_matrix(ctx, list->d_points, d_rp, d_col, d_cf, live, n, b->d_points, st)
                  : GroupOps<Fp2>::list_mul_matrix(ctx, list->d_points, d_rp, d_col, d_cf, live, n, b->d_points, st);
    if (rcb) return rcb;
    int rc = register_points(ctx, b, st);
    if (rc) return rc;
    CK(cudaStreamSynchronize(st));          // the temporaries are in use until here
    *out = bg.release();
    return BMPC_OK;
}

int bmpc_fixed_base_mul(bmpc_ctx* ctx, int group, const uint8_t* base, const uint64_t* scalars, size_t n,
                        int scalars_on_device, bmpc_bases** out) {
    if (!ctx || !base || (!scalars && n) || !out || (group != BMPC_G1 && group != BMPC_G2)) return BMPC_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = ctx->own_stream;
    StreamScope ss(ctx, st);
    size_t pb = group == BMPC_G1 ? 96 : 192;
    size_t xb = group == BMPC_G1 ? sizeof(G1XYZZ) : sizeof(G2XYZZ);
    DevBuf raw_buf, base_buf, table_buf, s_buf;
    BasesGuard bg(ctx);
    CK(cudaMalloc(&raw_buf.p, pb));
    CK(cudaMalloc(&base_buf.p, pb));
    CK(cudaMalloc(&table_buf.p, 32 * 256 * xb));
    uint8_t* d_base_raw = raw_buf.as<uint8_t>();
    void* d_base = base_buf.p;
    void* d_table = table_buf.p;
    CK(cudaMemcpyAsync(d_base_raw, base, pb, cudaMemcpyHostToDevice, st));
    const uint32_t* sc = (const uint32_t*)scalars;
    if (!scalars_on_device && n) {
        CK(cudaMalloc(&s_buf.p, n * 32));
        CK(cudaMemcpyAsync(s_buf.p, scalars, n * 32, cudaMemcpyHostToDevice, st));
        sc = s_buf.as<uint32_t>();
    }
    bmpc_bases* b = bg.b = new bmpc_bases();
    b->group = group;
    b->n = n;
    CK(cudaMalloc(&b->d_points, (n ? n : 1) * pb));
    int rcf;
    if (group == BMPC_G1) {
        rcf = GroupOps<Fp>::decode(ctx, d_base_raw, pb, 1, d_base, st);
        if (!rcf) rcf = GroupOps<Fp>::fixed_base_mul(ctx, d_base, d_table, sc, n, b->d_points, st);
    } else {
        rcf = GroupOps<Fp2>::decode(ctx, d_base_raw, pb, 1, d_base, st);
        if (!rcf) rcf = GroupOps<Fp2>::fixed_base_mul(ctx, d_base, d_table, sc, n, b->d_points, st);
    }
    if (rcf) return rcf;
    int rc = register_points(ctx, b, st);
    if (rc) return rc;
    CK(cudaStreamSynchronize(st));
    *out = bg.release();
    return BMPC_OK;
}

}  // extern "C"
