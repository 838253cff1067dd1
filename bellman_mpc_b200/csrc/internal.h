// Internal host-side declarations shared by the translation units of libbellman_b200.so.
// (api.cu = C ABI, ntt.cu = EvaluationDomain kernels, msm_sort.cu = digit/sort kernels,
//  group_g1.cu / group_g2.cu = curve kernels instantiated per group, prove.cu = proof tail.)
#pragma once
#include "../../include/bellman_b200.h"

#include <cuda_runtime.h>

#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "curve.cuh"

namespace bmpc {

constexpr uint32_t SMALL_LOG = 12;   // largest in-block NTT radix 2^12 (128 KB of shared memory)
constexpr uint32_t DIRECT_TABLE_MAX_LOG = 24;  // fully expanded power tables up to this domain size

enum TableKind {
    K_TW_FWD = 0, K_TW_INV, K_G, K_G_MINV, K_GINV_MINV, K_GINV_MINV_ZINV_CANON, K_COUNT
};

struct DevTable {
    Fr* hi = nullptr;
    Fr* lo = nullptr;
    Fr* direct = nullptr;
    uint32_t lo_bits = 0, hi_n = 0;
    bool ready = false;
};

struct DomainTables {
    Fr* d_consts = nullptr;  // 10 x Fr, see domain_consts_kernel
    Fr h_consts[10];
    DevTable t[K_COUNT];
};

}  // namespace bmpc

// Tuning knobs read from BMPC_* environment variables ONCE per context (bmpc_ctx_create) and again on
// bmpc_ctx_reload_env -- not per call.  -1 / 0 = automatic unless noted.
struct bmpc_tuning {
    int acc_pairs = 0;             // BMPC_ACC_PAIRS: 0 never (default), 1 always, -1 from pair_min_entries up (msm_pairs.cuh)
    size_t pair_min_entries = (size_t)1 << 22;   // BMPC_PAIR_MIN_ENTRIES
    int pair_k = 0;                // BMPC_PAIR_K: pairs per thread per inversion
    int acc_affine = -1;           // BMPC_ACC_AFFINE: 0 never, 1 always the batched-affine tree
    int aff_blockdim = 0, aff_ksel = 0, aff_minb = 0, aff_gmax = 0, aff_waves = 0, aff_force_g = 0;
    int aff_whole_waves = 1;
    int acc_compact = 0;           // BMPC_ACC_COMPACT
    int subwindows = 1;            // BMPC_MSM_SUBWINDOWS
    int reduce_block = 0;          // BMPC_REDUCE_BLOCK
    int reduce_slog_add = 0;       // BMPC_REDUCE_SLOG_ADD: more buckets per reduce thread than one wave needs (x 2^k)
    int ntt_no_direct = 0;         // BMPC_NTT_NO_DIRECT
    int sort_radix = -1;           // BMPC_SORT_RADIX: two-level partition sort, 0 never, 1 whenever possible, -1 automatic
    int rs_chunk_log = 0;          // BMPC_RS_CHUNK_LOG: entries per block of rs_bucket_hist / rs_scatter (0: 2^12)
    int tail_quad = 1;             // BMPC_TAIL_QUAD: fold / final steps of the bucket reduction with four lanes per element (G1)
    int proof_slots = 0;           // BMPC_PROOF_SLOTS: 3 chains of multiexps in create_proof, 8 (one stream each), 0 auto
    void load();
};

struct bmpc_ctx {
    int device = 0;
    bmpc_tuning tune;
    std::string err;
    std::mutex mu;
    uint64_t launches = 0;
    int tune_c = 0, tune_maxdeg = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // H2D of large inputs overlapping compute (create_proof)
    cudaEvent_t copy_done = nullptr;
    // scratch arena (grown on demand, reused across calls)
    char* ws = nullptr;
    size_t ws_size = 0, ws_used = 0;
    // extra multiexp slots (create_proof runs three chains of multiexps side by side): each has its
    // own stream, scratch arena and device staging words; slot 0 is the context's own
    struct Slot {
        cudaStream_t stream = nullptr;
        char* ws = nullptr;
        size_t ws_size = 0;
        uint8_t* d_stage = nullptr;
    };
    Slot slots[7];
    cudaEvent_t inputs_ready = nullptr;
    // Stream contract: the scratch arena, the twiddle / power tables and the staging words are shared
    // per context, so calls on DIFFERENT streams must not overlap.  Every stream-taking entry point
    // opens a StreamScope: a call arriving on another stream than the previous one first waits (on
    // the device) for an event recorded at the end of that previous call.
    cudaStream_t last_stream = nullptr;
    cudaEvent_t last_done = nullptr;
    bool last_valid = false;
    // device staging for host-pointer entry points (scalars, density words); grow-only so that a
    // prover calling in a loop does not pay cudaMalloc/cudaFree per call
    char* io = nullptr;
    size_t io_size = 0;
    // staging for small results
    uint8_t* h_stage = nullptr;  // 4 KB pinned
    uint8_t* d_stage = nullptr;  // 4 KB
    // Number of dense positions of the NEXT multiexp enqueued on this context, when its caller has the
    // density words on the host (exact popcount; 0 = unknown).  The plan then sizes the slices, the
    // scratch strides and the entry arrays for the points that really arrive instead of for one per
    // position (create_proof: B density 0.5 -> slices and scratch strides twice too long).  Consumed
    // (reset) by multiexp_enqueue; set under the context's lock.
    size_t dense_hint = 0;
    // lanes of the asynchronous multiexp (multi.cu: bmpc_multiexp_async): child contexts on the same
    // device, each with its own stream, scratch arena and staging, handed out one per Waiter
    std::vector<bmpc_ctx*> lanes;
    std::vector<char> lane_busy;
    std::mutex lanes_mu;
    std::map<uint32_t, bmpc::DomainTables> domains;
    bmpc::Fr* tw_small[2] = {nullptr, nullptr};  // fwd / inv powers of the 2^SMALL_LOG-th root
    // optional per-kernel timing (bench.py roofline): CUDA event pairs on the launching stream
    bool profile = false;
    struct ProfEvent { int id; cudaEvent_t e0, e1; };
    std::vector<ProfEvent> prof_pending;
    double prof_ms[BMPC_PROF_COUNT] = {0};
    uint64_t prof_n[BMPC_PROF_COUNT] = {0};
};

struct bmpc_bases {
    int group = 0;
    size_t n = 0;
    void* d_points = nullptr;   // Affine<Fp>[n] or Affine<Fp2>[n], Montgomery; with precomputed
                                // window tables: [tab_W][n], table w holding 2^(tab_c w) * P_i
    uint32_t* d_inf = nullptr;  // identity bitmap
    uint32_t tab_c = 0, tab_W = 0;  // 0 = no tables
};

struct bmpc_domain {
    bmpc::Fr* d = nullptr;
    size_t m = 0;
    uint32_t exp = 0;
};

namespace bmpc {

#define CK(call)                                                                      \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) {                                                      \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);            \
            return BMPC_ERR_CUDA;                                                     \
        }                                                                             \
    } while (0)

// BMPC_SYNC_CHECK=1 synchronises after every launch so that a fault is reported at its kernel.
inline bool bmpc_sync_check() {
    static const bool on = getenv("BMPC_SYNC_CHECK") && atoi(getenv("BMPC_SYNC_CHECK")) != 0;
    return on;
}

#define LAUNCH(ctx, kernel, grid, block, smem, stream, ...)                           \
    do {                                                                              \
        kernel<<<grid, block, smem, stream>>>(__VA_ARGS__);                           \
        (ctx)->launches++;                                                            \
        cudaError_t e_ = cudaGetLastError();                                          \
        if (e_ == cudaSuccess && bmpc::bmpc_sync_check()) e_ = cudaDeviceSynchronize(); \
        if (e_ != cudaSuccess) {                                                      \
            (ctx)->err = std::string(#kernel) + ": " + cudaGetErrorString(e_);        \
            return BMPC_ERR_CUDA;                                                     \
        }                                                                             \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// Orders a call on `st` after the context's previous call if that ran on another stream, and
// publishes its own completion point on exit (see bmpc_ctx::last_done).
struct StreamScope {
    bmpc_ctx* ctx;
    cudaStream_t st;
    StreamScope(bmpc_ctx* c, cudaStream_t s) : ctx(c), st(s) {
        if (!ctx->last_done) cudaEventCreateWithFlags(&ctx->last_done, cudaEventDisableTiming);
        if (ctx->last_valid && ctx->last_stream != st) cudaStreamWaitEvent(st, ctx->last_done, 0);
    }
    ~StreamScope() {
        if (ctx->last_done && cudaEventRecord(ctx->last_done, st) == cudaSuccess) {
            ctx->last_stream = st;
            ctx->last_valid = true;
        }
    }
};

// device / handle temporaries of the setup-time entry points: released on every exit path
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

// RAII event pair around one kernel launch (only when profiling is on)
struct ProfScope {
    bmpc_ctx* ctx;
    cudaStream_t st;
    bmpc_ctx::ProfEvent ev;
    bool on;
    ProfScope(bmpc_ctx* c, int id, cudaStream_t s) : ctx(c), st(s), on(c->profile) {
        if (!on) return;
        ev.id = id;
        cudaEventCreate(&ev.e0);
        cudaEventCreate(&ev.e1);
        cudaEventRecord(ev.e0, st);
    }
    ~ProfScope() {
        if (!on) return;
        cudaEventRecord(ev.e1, st);
        ctx->prof_pending.push_back(ev);
    }
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline size_t ws_need(size_t count, size_t elem) { return align_up(count * elem, 256); }

// Arena: reserve once per call with the total, then carve.
inline int ws_reserve(bmpc_ctx* ctx, size_t bytes) {
    ctx->ws_used = 0;
    if (bytes <= ctx->ws_size) return BMPC_OK;
    CK(cudaDeviceSynchronize());
    if (ctx->ws) CK(cudaFree(ctx->ws));
    ctx->ws = nullptr;
    ctx->ws_size = 0;
    size_t want = align_up(bytes + (bytes >> 3), 1 << 20);
    CK(cudaMalloc(&ctx->ws, want));
    ctx->ws_size = want;
    return BMPC_OK;
}
inline int io_reserve(bmpc_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->io_size) return BMPC_OK;
    CK(cudaDeviceSynchronize());
    if (ctx->io) CK(cudaFree(ctx->io));
    ctx->io = nullptr;
    ctx->io_size = 0;
    size_t want = align_up(bytes, 1 << 20);
    CK(cudaMalloc(&ctx->io, want));
    ctx->io_size = want;
    return BMPC_OK;
}
template <class T>
inline T* ws_take(bmpc_ctx* ctx, size_t count) {
    size_t bytes = align_up(count * sizeof(T), 256);
    if (ctx->ws_used + bytes > ctx->ws_size) return nullptr;
    T* p = reinterpret_cast<T*>(ctx->ws + ctx->ws_used);
    ctx->ws_used += bytes;
    return p;
}
inline cudaStream_t pick_stream(bmpc_ctx* ctx, void* stream) {
    return stream ? reinterpret_cast<cudaStream_t>(stream) : ctx->own_stream;
}

// ---------------------------------------------------------------- ntt.cu
int ntt_dev_locked(bmpc_ctx* ctx, Fr* d, uint32_t logm, int op, cudaStream_t st);
int h_coefficients_locked(bmpc_ctx* ctx, Fr* a, Fr* b, Fr* c, uint32_t logm, Fr* t1, Fr* t2,
                          cudaStream_t st);
int h_coset_evals_locked(bmpc_ctx* ctx, Fr* p, uint32_t logm, Fr* t1, Fr* t2, cudaStream_t st);
int h_from_coset_evals_locked(bmpc_ctx* ctx, Fr* a, const Fr* b, const Fr* c, uint32_t logm, Fr* t1, Fr* t2,
                              cudaStream_t st);
int fr_pointwise(bmpc_ctx* ctx, int what, Fr* a, const Fr* b, size_t n, cudaStream_t st);  // 0 mul 1 sub 2 to_canonical
int fr_scale_zinv(bmpc_ctx* ctx, Fr* a, size_t m, uint32_t logm, cudaStream_t st);
int fr_distribute_powers(bmpc_ctx* ctx, Fr* a, size_t m, const Fr* d_g, cudaStream_t st);
int fr_eval_z(bmpc_ctx* ctx, const Fr* d_tau, uint32_t logm, Fr* d_out, cudaStream_t st);
void ntt_free_tables(bmpc_ctx* ctx);
// pieces of the distributed four-step transform
int ntt_batch_dev_locked(bmpc_ctx* ctx, Fr* d, uint32_t logn, uint32_t batch, bool inverse, cudaStream_t st);
int fr_swap01(bmpc_ctx* ctx, const Fr* in, Fr* out, uint32_t d0, uint32_t d1, uint32_t d2, cudaStream_t st);
int fr_fourstep_twiddle(bmpc_ctx* ctx, Fr* d, uint32_t rows, uint32_t cols, uint32_t row0, uint32_t logm,
                        bool inverse, cudaStream_t st);
int fr_scale_pow(bmpc_ctx* ctx, Fr* d, size_t n, uint32_t first, uint32_t logm, int which, cudaStream_t st);

// ------------------------------------------------------------- msm_sort.cu
enum { MSM_FLAG_EOF = 1, MSM_FLAG_IDENT_ANY = 2, MSM_FLAG_IDENT_TOP = 4 };
// a bucket with at most this many partial sums is folded inline by the reduce kernel; more go
// through the heavy-bucket combine first (which leaves the total in the first slot)
#define BMPC_INLINE_PARTIALS 4u

// msm_fold_kernel: block results of the bucket reduction are folded in groups of this many before
// the final kernel; a set may have up to BMPC_FOLD_GROUP * 256 reduce blocks
#define BMPC_FOLD_GROUP 32u
struct MsmGeom {
    uint32_t c;        // window bits
    uint32_t W;        // number of windows = 255 / c + 1
    uint32_t B;        // buckets per window = 2^(c-1)  (signed digits)
    uint32_t L;        // max points per accumulate task
    uint32_t H;        // bucket sets: W without precomputed tables, 1 with them
    uint32_t tab_stride; // entries per precomputed table (0 without tables)
    uint32_t c_ref;    // the reference's window size for this n (error precedence only)
    uint32_t top_skip; // bit offset of the reference's highest window
};

struct MsmPlan {
    MsmGeom g;
    uint32_t nb;  // total buckets
    size_t max_pairs, max_tasks;
    // weighted-sum (bucket reduction) geometry per bucket set: each thread owns 2^s_log consecutive
    // buckets, tpw = B >> s_log threads per set in nblk (<= 8192) blocks of rblock (<= 256) threads
    uint32_t s_log, tpw, rblock, nblk;
    // batched-affine accumulation (msm_affine.cuh): decided per group in GroupOps::plan_affine
    bool affine = false, aff_whole_waves = false;
    uint32_t aff_G = 0, aff_blocks = 0, aff_block = 128, aff_K = 128, aff_minb = 1, aff_HA = 0, aff_HB = 0;
    // round-based pair accumulation (msm_pairs.cuh): decided per group in GroupOps::plan_affine.
    // pair_R rounds (slices of at most 2^pair_R entries, every bucket padded to an even count),
    // pair_out[r] / pair_list[r]: first pool index / first list entry of round r (host-side bounds)
    bool pairs = false;
    uint32_t pair_R = 0, pair_blocks = 0, pair_block = 128, pair_minb = 3, pair_kmax = 256;
    uint32_t pair_out[16] = {0}, pair_list[16] = {0};
    uint32_t pair_stride = 0, pair_cstride = 0;   // row pitch of the per-round task scans / their chunk sums
    size_t n = 0;
    bool has_density = false;
    // two-level partition sort instead of the one-pass scatter (msm_sort_kernels.cuh: rs_*); radix_ok:
    // geometry and size allow it, radix: chosen (never together with pair rounds, whose padding it lacks)
    bool radix_ok = false, radix = false;
    size_t sort_bytes;  // scratch for everything except the curve-typed buffers
};
// (re)derives max_tasks and sort_bytes from g.L / pairs (called by msm_make_plan and again by
// plan_affine when it switches the plan to pair rounds)
void msm_plan_sizes(MsmPlan& p);

struct MsmSorted {      // outputs of the sort stage (device pointers into the arena)
    uint32_t* sorted;   // base index | sign << 31, grouped by bucket
    uint32_t* off;      // nb + 1 bucket offsets
    uint32_t* toff;     // nb + 1 task offsets
    uint32_t* heavy;    // nb + 1 scratch for the heavy-bucket list
    uint32_t* heavy_count;
    uint4* desc;        // per task {first sorted entry, length, partial slot, bucket}, big tasks first
    uint32_t* ntasks;   // device pointer to the task count (== toff[nb])
    uint32_t* nsorted;  // device pointer to the number of entries in `sorted` (== off[nb]; padded in pair mode)
};

MsmPlan msm_make_plan(bmpc_ctx* ctx, const bmpc_bases* bases, size_t n, bool has_density, size_t n_ref = 0,
                      size_t n_dense = 0);
// set bits among the first nbits of a bitvec's raw words (host)
inline size_t popcount_bits(const uint64_t* words, size_t nbits) {
    size_t full = nbits / 64, cnt = 0;
    for (size_t i = 0; i < full; i++) cnt += (size_t)__builtin_popcountll(words[i]);
    if (nbits % 64) cnt += (size_t)__builtin_popcountll(words[full] & ((1ull << (nbits % 64)) - 1));
    return cnt;
}
uint32_t msm_table_window(size_t n_bases);  // window bits used for precomputed tables
// count -> scan -> scatter; raises EOF / identity flags into d_flags[0]
int msm_sort_run(bmpc_ctx* ctx, const MsmPlan& p, const bmpc_bases* bases, size_t base_offset,
                 const uint32_t* d_scalars, size_t n, const uint32_t* d_density, uint32_t* d_flags,
                 MsmSorted* out, cudaStream_t st);

// pair mode (msm_pairs.cuh): scans the tasks' pairs per round and writes the lists of rounds
// 1 .. R-1 (uint2 pool indices) and every task's result index `fin`
int msm_pairs_prepare(bmpc_ctx* ctx, const MsmPlan& p, const MsmSorted& s, uint32_t* csums, uint32_t* totals,
                      uint32_t* pairoff, uint2* lists, uint32_t* fin, cudaStream_t st);

// -------------------------------------------------- group_g1.cu / group_g2.cu
template <class F>
struct GroupOps {
    // chooses between the XYZZ and the batched-affine accumulate kernel (fills p.affine, p.aff_*)
    static void plan_affine(bmpc_ctx* ctx, MsmPlan& p);
    static size_t curve_bytes(const MsmPlan& p);
    static size_t curve_bytes_xyzz(const MsmPlan& p);
    // accumulate -> combine -> reduce -> final.  mode 0: uncompressed bytes to d_out_bytes;
    // mode 1: XYZZ partial to d_out_xyzz.
    static int msm_finish(bmpc_ctx* ctx, const MsmPlan& p, const bmpc_bases* bases, const MsmSorted& s,
                          int mode, uint8_t* d_out_bytes, void* d_out_xyzz, cudaStream_t st);
    static int sum_partials(bmpc_ctx* ctx, const void* d_parts, uint32_t count, uint8_t* d_out_bytes,
                            cudaStream_t st);
    static int decode(bmpc_ctx* ctx, const uint8_t* d_raw, size_t stride, size_t n, void* d_points,
                      cudaStream_t st);
    static int encode(bmpc_ctx* ctx, const void* d_points, size_t n, uint8_t* d_out, cudaStream_t st);
    // validating decode; *d_err must hold 0xffffffff on entry (atomicMin of index*4+kind)
    static int validate_decode(bmpc_ctx* ctx, const uint8_t* d_raw, size_t stride, size_t n, int checked,
                               int reject_identity, void* d_points, uint32_t* d_err, cudaStream_t st);
    static int inf_bitmap(bmpc_ctx* ctx, const void* d_points, size_t n, uint32_t* d_bitmap, cudaStream_t st);
    static int batch_mul(bmpc_ctx* ctx, const void* d_in, const uint32_t* d_scalars, int per_element,
                         size_t n, void* d_out, cudaStream_t st);
    // mpc.rs:416-457: out[i] = sum_j coeff[j] * list[col[j]] over CSR row i (i < live_rows), identity after
    static int list_mul_matrix(bmpc_ctx* ctx, const void* d_list, const uint32_t* d_row_ptr, const uint32_t* d_col,
                               const uint32_t* d_coeffs, size_t live_rows, size_t n_out, void* d_out,
                               cudaStream_t st);
    static int fixed_base_mul(bmpc_ctx* ctx, const void* d_base, void* d_table, const uint32_t* d_scalars,
                              size_t n, void* d_out, cudaStream_t st);
    // d_tables: [W][n] with table 0 already filled; fills tables 1..W-1 (2^(c w) * P_i, affine)
    static int precompute_tables(bmpc_ctx* ctx, void* d_tables, size_t n, uint32_t c, uint32_t W,
                                 cudaStream_t st);
};

// ------------------------------------------------------------------ api.cu
// A multiexp in flight: everything is enqueued on `st`, the flags word (and the result bytes)
// land in the pinned staging area at `h`; multiexp_collect turns them into the reference's status
// once the stream has been synchronised.  create_proof enqueues all eight before waiting once; the
// asynchronous Waiter of multi.cu keeps one per lane.
struct MsmPending {
    cudaStream_t st = nullptr;     // nullptr: nothing was launched (n == 0), status OK
    const uint8_t* h = nullptr;
    size_t out_bytes = 0;
    uint8_t* out = nullptr;
};
int flags_to_status(uint32_t flags);
// n_ref: the length of the WHOLE exponent vector when this call is one shard of it (0: this call is
// the whole multiexp)
int multiexp_enqueue(bmpc_ctx* ctx, const bmpc_bases* bases, size_t base_offset, const uint64_t* d_scalars,
                     size_t n, const uint64_t* d_density, size_t density_len, uint8_t* out, void* d_partial,
                     cudaStream_t st, uint8_t* h_dst, MsmPending* pend, size_t n_ref = 0);
int multiexp_collect(const MsmPending& pend, uint32_t* flags_out = nullptr);   // after the stream was synchronised
int multiexp_dev_locked(bmpc_ctx* ctx, const bmpc_bases* bases, size_t base_offset, const uint64_t* d_scalars,
                        size_t n, const uint64_t* d_density, size_t density_len, uint8_t* out, void* d_partial,
                        cudaStream_t st, size_t n_ref = 0, uint32_t* flags_out = nullptr);

// One device's share of create_proof (multi.cu; the eight multiexps in the order
// a_inputs, a_aux, b_g1_inputs, b_g1_aux, b_g2_inputs, b_g2_aux, h, l): multiexp j runs over the
// exponent positions [lo[j], hi[j]) of the FULL input / aux / H vector, starts at base
// base_offset[j] of the device's slice of the query vector and reads density words re-based to bit
// 0 of lo[j] (host memory, NULL = FullDensity); n_total[j] = length of the whole exponent vector.
struct ProofSlices {
    size_t lo[8], hi[8], base_offset[8], n_total[8];
    const uint64_t* dens[8];
};
int create_proof_common(bmpc_ctx* ctx, const bmpc_params* P, const bmpc_assignment* S, const uint64_t r[4],
                        const uint64_t s[4], const ProofSlices* slices, uint8_t* proof_out, uint8_t* partials_out,
                        uint32_t* flags_out);

// ---------------------------------------------------------------- prove.cu
struct ProveTailArgs {
    // multiexp partial sums (XYZZ), in the order the reference awaits them (prover.rs:328-343)
    const G1XYZZ* a_inputs; const G1XYZZ* a_aux;
    const G1XYZZ* b1_inputs; const G1XYZZ* b1_aux;
    const G2XYZZ* b2_inputs; const G2XYZZ* b2_aux;
    const G1XYZZ* h; const G1XYZZ* l;
    // vk: alpha_g1, beta_g1, delta_g1 (G1Affine[3]); beta_g2, delta_g2 (G2Affine[2])
    const G1Affine* vk_g1; const G2Affine* vk_g2;
    const Fr* rs;   // r, s in Montgomery form
    uint8_t* proof; // 192 B
};
int prove_tail_launch(bmpc_ctx* ctx, const ProveTailArgs& args, cudaStream_t st);
int prove_fold_partials_launch(bmpc_ctx* ctx, const uint8_t* d_all, uint32_t world, G1XYZZ* out1, G2XYZZ* out2,
                               cudaStream_t st);

}  // namespace bmpc
