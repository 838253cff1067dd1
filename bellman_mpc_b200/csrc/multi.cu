// Asynchronous multiexp (the reference's Waiter, src/multicore.rs:33-118) and the multi-device
// context: N GPUs of one node driven from ONE process through one C call per multiexp / proof
// (SURVEY 8b `bmpc_ctx_create(devices, n)`, 8e).  The reference's `multiexp` (multiexp.rs:254-281) and
// `create_proof` (groth16/prover.rs:176-350) are single calls in a single process; a Rust host keeps
// that shape and the library owns the sharding plan, the gather of the partial sums over NVLink
// (cudaMemcpyPeerAsync, 192 / 384 bytes per device) and the reference's error precedence across
// shards (multiexp.rs:244-249).  No field or group arithmetic here: host code only moves bytes,
// counts density bits and calls the per-device entry points of api.cu.
#include <algorithm>
#include <cstring>
#include <thread>

#include "internal.h"

using namespace bmpc;

// ------------------------------------------------------------------------------ async multiexp
struct bmpc_waiter {
    bmpc_ctx* parent = nullptr;
    int lane = -1;                 // index into parent->lanes; -1: finished at enqueue (n == 0 / error)
    MsmPending pend;
    int early_status = BMPC_OK;
    int group = BMPC_G1;
};

namespace {

constexpr size_t MAX_LANES = 16;

int lane_acquire(bmpc_ctx* ctx, int* lane_out) {
    std::lock_guard<std::mutex> lk(ctx->lanes_mu);
    for (size_t i = 0; i < ctx->lanes.size(); i++)
        if (!ctx->lane_busy[i]) {
            ctx->lane_busy[i] = 1;
            *lane_out = (int)i;
            return BMPC_OK;
        }
    if (ctx->lanes.size() >= MAX_LANES) {
        ctx->err = "too many multiexps in flight on one context (16): wait() on an earlier Waiter first";
        return BMPC_ERR_INVALID;
    }
    bmpc_ctx* lane = nullptr;
    int rc = bmpc_ctx_create(ctx->device, &lane);
    if (rc) return rc;
    lane->tune = ctx->tune;
    lane->tune_c = ctx->tune_c;
    ctx->lanes.push_back(lane);
    ctx->lane_busy.push_back(1);
    *lane_out = (int)ctx->lanes.size() - 1;
    return BMPC_OK;
}

void lane_release(bmpc_ctx* ctx, int lane) {
    std::lock_guard<std::mutex> lk(ctx->lanes_mu);
    ctx->lane_busy[lane] = 0;
}

// ---- density bit helpers (host) ---------------------------------------------------------------
inline size_t popcount_below(const uint64_t* words, size_t nbits) { return popcount_bits(words, nbits); }
// position of the k-th (0-based) set bit among the first nbits, or nbits if there are not that many;
// words == NULL: every position is dense
size_t dense_select(const uint64_t* words, size_t nbits, size_t k) {
    if (!words) return k < nbits ? k : nbits;
    const size_t nw = (nbits + 63) / 64;
    for (size_t i = 0; i < nw; i++) {
        uint64_t w = words[i];
        if (i == nw - 1 && nbits % 64) w &= (1ull << (nbits % 64)) - 1;
        const size_t c = (size_t)__builtin_popcountll(w);
        if (k < c) {
            for (;;) {
                const int b = __builtin_ctzll(w);
                if (k == 0) return i * 64 + (size_t)b;
                w &= w - 1;
                k--;
            }
        }
        k -= c;
    }
    return nbits;
}
// bits [lo, hi) re-based to bit 0
std::vector<uint64_t> slice_bits(const uint64_t* words, size_t lo, size_t hi) {
    const size_t n = hi - lo, nw = (n + 63) / 64;
    std::vector<uint64_t> out(nw ? nw : 1, 0);
    const size_t w0 = lo / 64, sh = lo % 64, last = (hi + 63) / 64;
    for (size_t i = 0; i < nw; i++) {
        uint64_t v = words[w0 + i] >> sh;
        if (sh && w0 + i + 1 < last) v |= words[w0 + i + 1] << (64 - sh);
        out[i] = v;
    }
    if (n % 64) out[nw - 1] &= (1ull << (n % 64)) - 1;
    return out;
}

}  // namespace

struct bmpc_multi {
    std::vector<bmpc_ctx*> ctx;
    std::vector<void*> d_part;        // per device: room for one XYZZ partial (512 B)
    void* d_gather = nullptr;         // device 0: one partial per device
    // the H polynomial shared between the devices (bmpc_multi_create_proof): a second context per device
    // (its main one is busy with the other seven multiexps meanwhile) and buffers of m coefficients
    std::vector<bmpc_ctx*> hctx;
    std::vector<std::vector<void*>> d_ev;   // [device][slot]: slots 0..2 on device 0, slot 0 elsewhere
    size_t ev_bytes = 0;
    std::string err;
    std::mutex mu;
};

struct bmpc_multi_bases {
    int group = 0;
    size_t n = 0;
    std::vector<bmpc_bases*> part;    // device g holds bases [lo[g], lo[g+1])
    std::vector<size_t> lo;           // G + 1 cut points
};

namespace {

// One multiexp's exponent positions cut so that device g's positions consume exactly bases
// [lo[g], lo[g+1]) of the vector (multiexp.rs:53-86: the k-th dense position reads base
// base_offset + k).  cut[g] .. cut[g+1] = positions of device g; the last device also takes every
// position past the end of the bases, so an overrun (EOF) is seen where the bases end.
struct MsmCuts {
    std::vector<size_t> cut;          // G + 1
    std::vector<size_t> base_off;     // first base inside the device's slice
};
MsmCuts plan_cuts(const bmpc_multi_bases* mb, size_t base_offset, size_t n, const uint64_t* density) {
    const size_t G = mb->part.size();
    MsmCuts c;
    c.cut.assign(G + 1, 0);
    c.base_off.assign(G, 0);
    for (size_t g = 0; g < G; g++) {
        const size_t B = mb->lo[g];
        const size_t k_lo = (B > base_offset ? B : base_offset) - base_offset;
        c.cut[g] = g == 0 ? 0 : dense_select(density, n, k_lo);
        c.base_off[g] = (base_offset > B ? base_offset : B) - B;
    }
    c.cut[G] = n;
    // a device whose slice lies entirely below base_offset consumes nothing: its offset may exceed
    // its slice length only together with an empty position range
    for (size_t g = 0; g + 1 <= G - 1; g++)
        if (c.cut[g + 1] < c.cut[g]) c.cut[g + 1] = c.cut[g];
    return c;
}

}  // namespace

extern "C" {

// =========================================================================== async multiexp
int bmpc_multiexp_async(bmpc_ctx* ctx, const bmpc_bases* bases, size_t base_offset, const uint64_t* scalars,
                        size_t n, const uint64_t* density_words, size_t density_len, bmpc_waiter** out) {
    if (!ctx || !bases || !out || (!scalars && n)) return BMPC_ERR_INVALID;
    *out = nullptr;
    if (density_words && density_len != n) return BMPC_ERR_LENGTH_MISMATCH;       // multiexp.rs:273-278
    bmpc_waiter* w = new bmpc_waiter();
    w->parent = ctx;
    w->group = bases->group;
    if (n == 0) {                       // identity, no error (SURVEY 8a'/8): nothing to enqueue
        *out = w;
        return BMPC_OK;
    }
    int li = -1;
    int rc = lane_acquire(ctx, &li);
    if (rc) { delete w; return rc; }
    bmpc_ctx* lane = ctx->lanes[li];
    {
        std::lock_guard<std::mutex> lk(lane->mu);
        DeviceGuard dg(lane->device);
        cudaStream_t st = lane->own_stream;
        auto fail = [&](int code, const std::string& msg) {
            ctx->err = msg;
            cudaStreamSynchronize(st);
            lane_release(ctx, li);
            delete w;
            return code;
        };
        const size_t dw = (n + 63) / 64, sbytes = align_up(n * 32, 256);
        rc = io_reserve(lane, sbytes + dw * 8 + 256);
        if (rc) return fail(rc, lane->err);
        uint64_t* d_s = reinterpret_cast<uint64_t*>(lane->io);
        uint64_t* d_d = nullptr;
        if (cudaMemcpyAsync(d_s, scalars, n * 32, cudaMemcpyHostToDevice, st) != cudaSuccess)
            return fail(BMPC_ERR_CUDA, "bmpc_multiexp_async: scalar upload failed");
        if (density_words) {
            d_d = reinterpret_cast<uint64_t*>(lane->io + sbytes);
            if (cudaMemcpyAsync(d_d, density_words, dw * 8, cudaMemcpyHostToDevice, st) != cudaSuccess)
                return fail(BMPC_ERR_CUDA, "bmpc_multiexp_async: density upload failed");
        }
        // the result lands in the lane's pinned staging; the caller's buffer is filled by wait()
        lane->dense_hint = density_words ? popcount_bits(density_words, n) : 0;
        rc = multiexp_enqueue(lane, bases, base_offset, d_s, n, d_d, density_len, lane->h_stage + 2048, nullptr, st,
                              lane->h_stage, &w->pend);
        if (rc) return fail(rc, lane->err);
    }
    w->lane = li;
    *out = w;
    return BMPC_OK;
}

int bmpc_waiter_wait(bmpc_waiter* w, uint8_t* out) {
    if (!w || !out) return BMPC_ERR_INVALID;
    int status = w->early_status;
    if (w->lane < 0) {
        if (status == BMPC_OK) {
            memset(out, 0, w->group == BMPC_G1 ? 96 : 192);
            out[0] = 0x40;
        }
        delete w;
        return status;
    }
    bmpc_ctx* lane = w->parent->lanes[w->lane];
    {
        std::lock_guard<std::mutex> lk(lane->mu);
        DeviceGuard dg(lane->device);
        if (w->pend.st && cudaStreamSynchronize(w->pend.st) != cudaSuccess) {
            w->parent->err = "bmpc_waiter_wait: stream synchronisation failed";
            status = BMPC_ERR_CUDA;
        } else {
            w->pend.out = out;
            status = multiexp_collect(w->pend);
        }
    }
    lane_release(w->parent, w->lane);
    delete w;
    return status;
}

// =========================================================================== multi-device context
int bmpc_multi_create(const int* devices, int n, bmpc_multi** out) {
    if (!devices || n < 1 || n > 64 || !out) return BMPC_ERR_INVALID;
    *out = nullptr;
    bmpc_multi* m = new bmpc_multi();
    for (int g = 0; g < n; g++) {
        bmpc_ctx* c = nullptr;
        int rc = bmpc_ctx_create(devices[g], &c);
        if (rc) {
            for (bmpc_ctx* x : m->ctx) bmpc_ctx_destroy(x);
            delete m;
            return rc;
        }
        m->ctx.push_back(c);
    }
    for (int g = 0; g < n; g++) {
        DeviceGuard dg(devices[g]);
        void* p = nullptr;
        if (cudaMalloc(&p, 512) != cudaSuccess) { bmpc_multi_destroy(m); return BMPC_ERR_CUDA; }
        m->d_part.push_back(p);
        // direct NVLink access for the partial-sum gather where the topology offers it
        for (int h = 0; h < n; h++)
            if (devices[h] != devices[g]) {
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, devices[g], devices[h]) == cudaSuccess && can)
                    if (cudaDeviceEnablePeerAccess(devices[h], 0) != cudaSuccess) cudaGetLastError();
            }
    }
    {
        DeviceGuard dg(devices[0]);
        if (cudaMalloc(&m->d_gather, 512 * (size_t)n) != cudaSuccess) { bmpc_multi_destroy(m); return BMPC_ERR_CUDA; }
    }
    m->d_ev.assign((size_t)n, std::vector<void*>());
    *out = m;
    return BMPC_OK;
}

void bmpc_multi_destroy(bmpc_multi* m) {
    if (!m) return;
    for (size_t g = 0; g < m->ctx.size(); g++) {
        DeviceGuard dg(m->ctx[g]->device);
        if (g < m->d_part.size() && m->d_part[g]) cudaFree(m->d_part[g]);
        if (g == 0 && m->d_gather) cudaFree(m->d_gather);
    }
    for (size_t g = 0; g < m->d_ev.size(); g++) {
        DeviceGuard dg(m->ctx[g]->device);
        for (void* p : m->d_ev[g])
            if (p) cudaFree(p);
    }
    for (bmpc_ctx* c : m->hctx) bmpc_ctx_destroy(c);
    for (bmpc_ctx* c : m->ctx) bmpc_ctx_destroy(c);
    delete m;
}

int bmpc_multi_size(const bmpc_multi* m) { return m ? (int)m->ctx.size() : 0; }
bmpc_ctx* bmpc_multi_ctx(bmpc_multi* m, int rank) {
    return (m && rank >= 0 && (size_t)rank < m->ctx.size()) ? m->ctx[rank] : nullptr;
}
const char* bmpc_multi_last_error(const bmpc_multi* m) { return m ? m->err.c_str() : "null context"; }

// ---- bases split over the devices ------------------------------------------------------------
int bmpc_multi_bases_register(bmpc_multi* m, int group, const void* points, size_t n, size_t stride, int form,
                              bmpc_multi_bases** out) {
    if (!m || !out || (group != BMPC_G1 && group != BMPC_G2) || (!points && n)) return BMPC_ERR_INVALID;
    *out = nullptr;
    const size_t pb = group == BMPC_G1 ? 96 : 192;
    if (stride == 0) stride = pb;
    const size_t G = m->ctx.size();
    bmpc_multi_bases* mb = new bmpc_multi_bases();
    mb->group = group;
    mb->n = n;
    mb->lo.assign(G + 1, 0);
    for (size_t g = 0; g <= G; g++) mb->lo[g] = n / G * g + std::min(g, n % G);
    mb->part.assign(G, nullptr);
    std::vector<int> rcs(G, BMPC_OK);
    std::vector<std::thread> th;
    for (size_t g = 0; g < G; g++)
        th.emplace_back([&, g]() {
            const uint8_t* src = reinterpret_cast<const uint8_t*>(points) + mb->lo[g] * stride;
            rcs[g] = bmpc_bases_register(m->ctx[g], group, src, mb->lo[g + 1] - mb->lo[g], stride, form, &mb->part[g]);
        });
    for (auto& t : th) t.join();
    for (size_t g = 0; g < G; g++)
        if (rcs[g]) {
            m->err = bmpc_last_error(m->ctx[g]);
            int rc = rcs[g];
            bmpc_multi_bases_free(m, mb);
            return rc;
        }
    *out = mb;
    return BMPC_OK;
}

int bmpc_multi_bases_precompute(bmpc_multi* m, bmpc_multi_bases* mb, int window_bits) {
    if (!m || !mb) return BMPC_ERR_INVALID;
    const size_t G = m->ctx.size();
    std::vector<int> rcs(G, BMPC_OK);
    std::vector<std::thread> th;
    for (size_t g = 0; g < G; g++)
        th.emplace_back([&, g]() { rcs[g] = bmpc_bases_precompute(m->ctx[g], mb->part[g], window_bits); });
    for (auto& t : th) t.join();
    for (size_t g = 0; g < G; g++)
        if (rcs[g]) { m->err = bmpc_last_error(m->ctx[g]); return rcs[g]; }
    return BMPC_OK;
}

size_t bmpc_multi_bases_len(const bmpc_multi_bases* mb) { return mb ? mb->n : 0; }
const bmpc_bases* bmpc_multi_bases_part(const bmpc_multi_bases* mb, int rank, size_t* first) {
    if (!mb || rank < 0 || (size_t)rank >= mb->part.size()) return nullptr;
    if (first) *first = mb->lo[rank];
    return mb->part[rank];
}

void bmpc_multi_bases_free(bmpc_multi* m, bmpc_multi_bases* mb) {
    if (!mb) return;
    for (size_t g = 0; g < mb->part.size(); g++)
        if (mb->part[g]) bmpc_bases_free(m && g < m->ctx.size() ? m->ctx[g] : nullptr, mb->part[g]);
    delete mb;
}

// ---- multiexp on all devices (multiexp.rs:254-281 in one call) ----------------------------------
int bmpc_multi_multiexp(bmpc_multi* m, const bmpc_multi_bases* mb, size_t base_offset, const uint64_t* scalars,
                        size_t n, const uint64_t* density_words, size_t density_len, uint8_t* out) {
    if (!m || !mb || !out || (!scalars && n)) return BMPC_ERR_INVALID;
    if (density_words && density_len != n) return BMPC_ERR_LENGTH_MISMATCH;       // multiexp.rs:273-278
    std::lock_guard<std::mutex> lk(m->mu);
    const size_t G = m->ctx.size();
    const size_t pbytes = bmpc_partial_bytes(mb->group);
    const MsmCuts cuts = plan_cuts(mb, base_offset, n, density_words);
    std::vector<int> rcs(G, BMPC_OK);
    std::vector<uint32_t> flags(G, 0);
    auto shard = [&](size_t g) {
        bmpc_ctx* ctx = m->ctx[g];
        const size_t lo = cuts.cut[g], hi = cuts.cut[g + 1], ns = hi - lo;
        std::lock_guard<std::mutex> lg(ctx->mu);
        DeviceGuard dg(ctx->device);
        cudaStream_t st = ctx->own_stream;
        StreamScope ss(ctx, st);
        auto run = [&]() -> int {
            uint64_t* d_s = nullptr;
            uint64_t* d_d = nullptr;
            std::vector<uint64_t> dens;
            if (ns) {
                const size_t dw = (ns + 63) / 64, sbytes = align_up(ns * 32, 256);
                int rr = io_reserve(ctx, sbytes + dw * 8 + 256);
                if (rr) return rr;
                d_s = reinterpret_cast<uint64_t*>(ctx->io);
                CK(cudaMemcpyAsync(d_s, scalars + 4 * lo, ns * 32, cudaMemcpyHostToDevice, st));
                if (density_words) {
                    dens = slice_bits(density_words, lo, hi);
                    d_d = reinterpret_cast<uint64_t*>(ctx->io + sbytes);
                    CK(cudaMemcpyAsync(d_d, dens.data(), dw * 8, cudaMemcpyHostToDevice, st));
                }
            }
            ctx->dense_hint = (d_d && ns) ? popcount_bits(dens.data(), ns) : 0;
            int rc = multiexp_dev_locked(ctx, mb->part[g], cuts.base_off[g], d_s, ns, d_d, ns, nullptr, m->d_part[g], st,
                                         n, &flags[g]);
            // a shard's own verdict is not the multiexp's: the flag words of all shards are ORed below
            return (rc == BMPC_ERR_UNEXPECTED_EOF || rc == BMPC_ERR_UNEXPECTED_IDENTITY) ? BMPC_OK : rc;
        };
        rcs[g] = run();
    };
    if (G == 1) shard(0);
    else {
        std::vector<std::thread> th;
        for (size_t g = 0; g < G; g++) th.emplace_back(shard, g);
        for (auto& t : th) t.join();
    }
    uint32_t flags_or = 0;
    for (size_t g = 0; g < G; g++) {
        if (rcs[g]) { m->err = bmpc_last_error(m->ctx[g]); return rcs[g]; }
        flags_or |= flags[g];
    }
    const int status = flags_to_status(flags_or);         // multiexp.rs:244-249 across the shards
    if (status != BMPC_OK) return status;
    // gather the partial sums on device 0 (peer copies over NVLink) and fold them there
    bmpc_ctx* c0 = m->ctx[0];
    {
        std::lock_guard<std::mutex> l0(c0->mu);
        DeviceGuard dg(c0->device);
        bmpc_ctx* ctx = c0;
        cudaStream_t st = c0->own_stream;
        for (size_t g = 0; g < G; g++) {
            uint8_t* dst = reinterpret_cast<uint8_t*>(m->d_gather) + g * pbytes;
            if (m->ctx[g]->device == c0->device)
                CK(cudaMemcpyAsync(dst, m->d_part[g], pbytes, cudaMemcpyDeviceToDevice, st));
            else
                CK(cudaMemcpyPeerAsync(dst, c0->device, m->d_part[g], m->ctx[g]->device, pbytes, st));
        }
    }
    int rc = bmpc_sum_partials(c0, mb->group, m->d_gather, G, out, nullptr);
    if (rc) m->err = bmpc_last_error(c0);
    return rc;
}

// ---- create_proof on all devices (prover.rs:206-350 in one call) --------------------------------
int bmpc_multi_create_proof(bmpc_multi* m, const bmpc_multi_params* MP, const bmpc_assignment* S,
                            const uint64_t r[4], const uint64_t s[4], uint8_t proof_out[192]) {
    if (!m || !MP || !S || !r || !s || !proof_out) return BMPC_ERR_INVALID;
    if (!MP->h || !MP->l || !MP->a || !MP->b_g1 || !MP->b_g2) return BMPC_ERR_INVALID;
    std::lock_guard<std::mutex> lk(m->mu);
    const size_t G = m->ctx.size();
    const size_t nc = S->num_constraints, ni = S->num_inputs, na = S->num_aux;
    size_t mm = 1;
    uint32_t exp = 0;
    while (mm < nc) {
        mm *= 2;
        if (++exp >= 32) return BMPC_ERR_DEGREE_TOO_LARGE;
    }
    const size_t b_in_total = ni ? popcount_below(S->b_input_density, ni) : 0;
    // the eight multiexps: which vector, first base, exponent count, density (prover.rs:233,252-307)
    const bmpc_multi_bases* vec[8] = {MP->a, MP->a, MP->b_g1, MP->b_g1, MP->b_g2, MP->b_g2, MP->h, MP->l};
    const size_t boff[8] = {0, ni, 0, b_in_total, 0, b_in_total, 0, 0};
    const size_t cnt[8] = {ni, na, ni, na, ni, na, mm - 1, na};
    const uint64_t* dens[8] = {nullptr, S->a_aux_density, S->b_input_density, S->b_aux_density,
                               S->b_input_density, S->b_aux_density, nullptr, nullptr};
    MsmCuts cuts[8];
    for (int j = 0; j < 8; j++) {
        if (vec[j]->part.size() != G) return BMPC_ERR_INVALID;
        cuts[j] = plan_cuts(vec[j], boff[j], cnt[j], dens[j]);
    }
    std::vector<int> rcs(G, BMPC_OK);
    std::vector<uint8_t> partials(G * BMPC_PROOF_PARTIAL_BYTES);
    std::vector<uint32_t> flags(G * 8, 0);
    std::vector<bmpc_params> P(G);
    // With two or more devices the H polynomial (prover.rs:210-231) is computed ONCE between them
    // instead of on every device (each would upload all of a, b, c and run seven transforms before its
    // H multiexp can start): devices 0 .. min(G, 3) - 1 each take one of a, b, c (upload, ifft,
    // coset_fft), device 0 pulls the other two coset evaluations over NVLink, finishes the pipeline and
    // every device pulls its slice of the m - 1 scalars for its share of the H multiexp -- all on second
    // contexts, under the other seven multiexps.
    const bool share_h = G >= 2 && mm >= 2;
    auto share = [&](size_t g) {
        bmpc_params& p = P[g];
        p.h = MP->h->part[g]; p.l = MP->l->part[g]; p.a = MP->a->part[g];
        p.b_g1 = MP->b_g1->part[g]; p.b_g2 = MP->b_g2->part[g];
        memcpy(p.alpha_g1, MP->alpha_g1, 96); memcpy(p.beta_g1, MP->beta_g1, 96); memcpy(p.beta_g2, MP->beta_g2, 192);
        memcpy(p.delta_g1, MP->delta_g1, 96); memcpy(p.delta_g2, MP->delta_g2, 192);
        ProofSlices sl;
        std::vector<uint64_t> dbuf[8];
        for (int j = 0; j < 8; j++) {
            sl.lo[j] = cuts[j].cut[g];
            sl.hi[j] = cuts[j].cut[g + 1];
            if (j == 6 && share_h) sl.hi[j] = sl.lo[j];        // H comes from the shared pipeline below
            sl.base_offset[j] = cuts[j].base_off[g];
            sl.n_total[j] = cnt[j];
            sl.dens[j] = nullptr;
            if (dens[j] && sl.hi[j] > sl.lo[j]) {
                // jobs over the same words and positions share one buffer (b_g1 / b_g2)
                int same = -1;
                for (int k = 0; k < j; k++)
                    if (dens[k] == dens[j] && sl.lo[k] == sl.lo[j] && sl.hi[k] == sl.hi[j] && sl.dens[k]) same = k;
                if (same >= 0) sl.dens[j] = sl.dens[same];
                else {
                    dbuf[j] = slice_bits(dens[j], sl.lo[j], sl.hi[j]);
                    sl.dens[j] = dbuf[j].data();
                }
            }
        }
        rcs[g] = create_proof_common(m->ctx[g], &p, S, nullptr, nullptr, &sl, nullptr,
                                     partials.data() + g * BMPC_PROOF_PARTIAL_BYTES, flags.data() + 8 * g);
    };
    std::vector<std::thread> th;
    if (G == 1) share(0);
    else
        for (size_t g = 0; g < G; g++) th.emplace_back(share, g);

    std::vector<int> hrc(G, BMPC_OK);
    std::vector<uint32_t> hflags(G, 0);
    std::vector<uint8_t> hpart(G * sizeof(G1XYZZ), 0);
    if (share_h) {
        // second contexts and coefficient buffers, created on first use and kept
        int setup = BMPC_OK;
        while (m->hctx.size() < G && setup == BMPC_OK) {
            bmpc_ctx* c = nullptr;
            setup = bmpc_ctx_create(m->ctx[m->hctx.size()]->device, &c);
            if (setup == BMPC_OK) m->hctx.push_back(c);
        }
        const size_t need = mm * 32;
        if (setup == BMPC_OK && need > m->ev_bytes) {
            for (size_t g = 0; g < G && setup == BMPC_OK; g++) {
                DeviceGuard dg(m->ctx[g]->device);
                for (void* q : m->d_ev[g]) cudaFree(q);
                m->d_ev[g].assign(g == 0 ? 3 : 1, nullptr);
                for (auto& q : m->d_ev[g])
                    if (cudaMalloc(&q, need) != cudaSuccess) setup = BMPC_ERR_CUDA;
            }
            m->ev_bytes = setup == BMPC_OK ? need : 0;
        }
        if (setup != BMPC_OK) {
            for (auto& t : th) t.join();
            m->err = "bmpc_multi_create_proof: setting up the shared H pipeline failed";
            return setup;
        }
        const uint64_t* vec[3] = {S->a, S->b, S->c};
        const size_t owners = G < 3 ? G : 3;
        auto owner = [&](int k) { return (size_t)k % owners; };
        // where vector k's coset evaluations are produced: device 0 keeps slot k, the others slot 0
        auto ev_of = [&](int k) { const size_t o = owner(k); return o == 0 ? m->d_ev[0][k] : m->d_ev[o][0]; };
        {   // stage 1: one vector per owner device
            std::vector<std::thread> hb;
            for (size_t o = 0; o < owners; o++)
                hb.emplace_back([&, o]() {
                    bmpc_ctx* hc = m->hctx[o];
                    DeviceGuard dg(hc->device);
                    for (int k = 0; k < 3; k++) {
                        if (owner(k) != o || hrc[o]) continue;
                        hrc[o] = bmpc_h_coset_evals_dev(hc, (uint64_t*)ev_of(k), exp, vec[k], nc, nullptr);
                    }
                    if (!hrc[o] && cudaStreamSynchronize(hc->own_stream) != cudaSuccess) hrc[o] = BMPC_ERR_CUDA;
                });
            for (auto& t : hb) t.join();
        }
        bool h_ok = true;
        for (size_t o = 0; o < owners; o++) h_ok = h_ok && hrc[o] == BMPC_OK;
        if (h_ok) {   // stage 2: device 0 pulls b's and c's evaluations and finishes the pipeline
            bmpc_ctx* hc = m->hctx[0];
            DeviceGuard dg(hc->device);
            for (int k = 1; k < 3 && h_ok; k++) {
                const size_t o = owner(k);
                if (o == 0) continue;
                cudaError_t e = m->ctx[o]->device == hc->device
                                    ? cudaMemcpyAsync(m->d_ev[0][k], ev_of(k), need, cudaMemcpyDeviceToDevice, hc->own_stream)
                                    : cudaMemcpyPeerAsync(m->d_ev[0][k], hc->device, ev_of(k), m->ctx[o]->device, need,
                                                          hc->own_stream);
                h_ok = e == cudaSuccess;
            }
            if (h_ok)
                hrc[0] = bmpc_h_from_coset_evals_dev(hc, (uint64_t*)m->d_ev[0][0], (const uint64_t*)m->d_ev[0][1],
                                                     (const uint64_t*)m->d_ev[0][2], exp, nullptr);
            if (!h_ok || (!hrc[0] && cudaStreamSynchronize(hc->own_stream) != cudaSuccess)) hrc[0] = BMPC_ERR_CUDA;
            h_ok = hrc[0] == BMPC_OK;
        }
        if (h_ok) {   // stage 3: every device pulls its slice of the scalars and runs its share of the H multiexp
            std::vector<std::thread> hb;
            for (size_t g = 0; g < G; g++)
                hb.emplace_back([&, g]() {
                    bmpc_ctx* ctx = m->hctx[g];
                    std::lock_guard<std::mutex> lg(ctx->mu);
                    DeviceGuard dg(ctx->device);
                    cudaStream_t st = ctx->own_stream;
                    StreamScope ss(ctx, st);
                    auto run = [&]() -> int {
                        const size_t lo = cuts[6].cut[g], hi = cuts[6].cut[g + 1], ns = hi - lo;
                        // device 0 reads its slice in place; the others into their slot 0 (their own
                        // coset evaluations are no longer needed)
                        const uint64_t* d_sc = reinterpret_cast<const uint64_t*>(m->d_ev[0][0]) + 4 * lo;
                        if (g != 0 && ns) {
                            void* dst = m->d_ev[g][0];
                            if (m->ctx[g]->device == m->ctx[0]->device)
                                CK(cudaMemcpyAsync(dst, d_sc, ns * 32, cudaMemcpyDeviceToDevice, st));
                            else
                                CK(cudaMemcpyPeerAsync(dst, ctx->device, d_sc, m->ctx[0]->device, ns * 32, st));
                            d_sc = reinterpret_cast<const uint64_t*>(dst);
                        }
                        int rc = multiexp_dev_locked(ctx, MP->h->part[g], cuts[6].base_off[g], d_sc, ns, nullptr, 0, nullptr,
                                                     m->d_part[g], st, cnt[6], &hflags[g]);
                        if (rc != BMPC_OK && rc != BMPC_ERR_UNEXPECTED_EOF && rc != BMPC_ERR_UNEXPECTED_IDENTITY) return rc;
                        CK(cudaMemcpyAsync(ctx->h_stage + 1024, m->d_part[g], sizeof(G1XYZZ), cudaMemcpyDeviceToHost, st));
                        CK(cudaStreamSynchronize(st));
                        memcpy(hpart.data() + g * sizeof(G1XYZZ), ctx->h_stage + 1024, sizeof(G1XYZZ));
                        return BMPC_OK;
                    };
                    hrc[g] = run();
                });
            for (auto& t : hb) t.join();
        }
    }
    for (auto& t : th) t.join();
    if (share_h) {
        for (size_t g = 0; g < G; g++)
            if (hrc[g]) { m->err = bmpc_last_error(m->hctx[g]); return hrc[g]; }
        for (size_t g = 0; g < G; g++) {        // the H partial is entry 4 of the six G1 sums (a_in, a_aux, b1_in, b1_aux, h, l)
            memcpy(partials.data() + g * BMPC_PROOF_PARTIAL_BYTES + 4 * sizeof(G1XYZZ), hpart.data() + g * sizeof(G1XYZZ),
                   sizeof(G1XYZZ));
            flags[8 * g + 6] = hflags[g];
        }
    }
    for (size_t g = 0; g < G; g++)
        if (rcs[g]) { m->err = bmpc_last_error(m->ctx[g]); return rcs[g]; }
    // subversion check on delta first, then the multiexp statuses in the order the reference awaits
    // them (prover.rs:309-343), each from the OR of the devices' flag words
    if ((MP->delta_g1[0] & 0x40) || (MP->delta_g2[0] & 0x40)) return BMPC_ERR_UNEXPECTED_IDENTITY;
    for (int j = 0; j < 8; j++) {
        uint32_t f = 0;
        for (size_t g = 0; g < G; g++) f |= flags[8 * g + j];
        const int stj = flags_to_status(f);
        if (stj != BMPC_OK) return stj;
    }
    int rc = bmpc_create_proof_finish(m->ctx[0], &P[0], partials.data(), G, r, s, proof_out);
    if (rc) m->err = bmpc_last_error(m->ctx[0]);
    return rc;
}

}  // extern "C"
