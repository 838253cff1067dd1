/* bellman_b200.h -- C ABI of libbellman_b200.so
 *
 * B200-native (sm_100a) replacement for the data-parallel hot path of
 * doubiliu/bellman-mpc (a bellman 0.11.1 fork): Groth16 multiexp (MSM), the radix-2 Fr
 * EvaluationDomain transforms, the H-polynomial pipeline + create_proof orchestration, and
 * the ceremony's batch scalar multiplication.
 *
 * The reference has no FFI for this path (its only `extern "C"` precedent is the toy
 * exports at src/lib.rs:156-164,179-201 on a crate-type = ["dylib"], Cargo.toml:48-50), so
 * every entry point below cites the Rust function it replaces; INTEGRATION.md shows the
 * `extern "C"` block + build.rs a maintainer adds to src/multiexp.rs, src/domain.rs and
 * src/groth16/prover.rs.
 *
 * Conventions
 *   - plain pointers and sizes only; all arrays little-endian u64/u32 limbs as Rust holds them
 *   - scalars ("exponents"): canonical (non-Montgomery) 256-bit little-endian integers,
 *     4 x u64 each == ff::FieldBits<[u64; 4]>                        (src/multiexp.rs:162)
 *   - Fr coefficients: 4 x u64 Montgomery limbs, fully reduced == domain::Scalar<Fr>
 *                                                                     (src/domain.rs:22,230)
 *   - points in/out: ZCash uncompressed big-endian (G1 96 B, G2 192 B = x.c1|x.c0|y.c1|y.c0,
 *     bit 0x40 of byte 0 = infinity) unless a "_mont" form is requested
 *   - density maps: bitvec BitVec<Lsb0, usize> raw words, bit i = bit i%64 of word i/64
 *                                                                     (src/multiexp.rs:117-157)
 *   - every function returns a bmpc_status; there is NO CPU fallback: without a usable
 *     CUDA device calls fail with BMPC_ERR_CUDA
 *   - "_dev" variants take device pointers (inputs already resident in HBM) and a
 *     cudaStream_t passed as void* (NULL = the context's own stream); the plain variants take host
 *     pointers and include the host<->device copies
 *   - stream contract: a context's scratch arena, twiddle tables and staging words are shared by
 *     all its calls.  Calls are serialised by a mutex on the host; on the device a call issued on
 *     another stream than the context's previous call first waits for that call (event), so
 *     asynchronous "_dev" calls on different streams never overlap on the shared state.  Use one
 *     context per concurrent stream to overlap work.
 */
#ifndef BELLMAN_B200_H
#define BELLMAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bmpc_ctx bmpc_ctx;       /* one per process per GPU                         */
typedef struct bmpc_bases bmpc_bases;   /* device-resident base vector (Arc<Vec<Affine>>)   */
typedef struct bmpc_domain bmpc_domain; /* device-resident EvaluationDomain<Fr, Scalar<Fr>> */

/* SynthesisError variants this path can produce (src/lib.rs:355-370) */
typedef enum {
    BMPC_OK = 0,
    BMPC_ERR_UNEXPECTED_IDENTITY = 1,   /* SynthesisError::UnexpectedIdentity              */
    BMPC_ERR_UNEXPECTED_EOF = 2,        /* SynthesisError::IoError(UnexpectedEof)          */
    BMPC_ERR_DEGREE_TOO_LARGE = 3,      /* SynthesisError::PolynomialDegreeTooLarge        */
    BMPC_ERR_LENGTH_MISMATCH = 4,       /* the reference's assert!/assert_eq! panics       */
    BMPC_ERR_CUDA = 5,                  /* -> SynthesisError::IoError                      */
    BMPC_ERR_INVALID = 6,               /* malformed argument                              */
    BMPC_ERR_INVALID_DATA = 7           /* io::ErrorKind::InvalidData ("invalid G1", "point at infinity") */
} bmpc_status;

enum { BMPC_G1 = 1, BMPC_G2 = 2 };
/* point forms accepted by bmpc_bases_register */
enum {
    BMPC_FORM_UNCOMPRESSED_BE = 0,      /* G1Affine::to_uncompressed() bytes               */
    BMPC_FORM_MONT_XY = 1               /* raw Montgomery limbs x|y, (0,0) == infinity     */
};
/* EvaluationDomain transforms (src/domain.rs:81-125) */
enum { BMPC_FFT = 0, BMPC_IFFT = 1, BMPC_COSET_FFT = 2, BMPC_ICOSET_FFT = 3 };

/* ---- context ------------------------------------------------------------------------ */
/* replaces Worker::new() (src/multicore.rs:25-27): the execution resource handed to every call */
int  bmpc_ctx_create(int device, bmpc_ctx** out);
void bmpc_ctx_destroy(bmpc_ctx* ctx);
const char* bmpc_last_error(const bmpc_ctx* ctx);
/* tuning knobs (0 = automatic): MSM window bits, NTT max radix log2 */
int  bmpc_ctx_set_tuning(bmpc_ctx* ctx, int msm_window_bits, int ntt_max_deg);
/* BMPC_* environment knobs (kernel selection for tests and tuning sweeps) are parsed once when the
 * context is created; this re-reads them */
int  bmpc_ctx_reload_env(bmpc_ctx* ctx);
/* number of kernels launched by this context since creation (bench.py "gpu_launches") */
uint64_t bmpc_ctx_launch_count(const bmpc_ctx* ctx);

/* Per-kernel device timing for the roofline report: when enabled, the named kernels are
 * bracketed by CUDA events on their launching stream.  bmpc_ctx_profile_read synchronises,
 * returns the accumulated milliseconds / launch count for `which` and resets them. */
enum { BMPC_PROF_MSM_ACCUMULATE = 0, BMPC_PROF_NTT_PASS = 1, BMPC_PROF_MSM_SORT = 2,
       BMPC_PROF_MSM_REDUCE = 3, BMPC_PROF_COUNT = 4 };
int  bmpc_ctx_profile(bmpc_ctx* ctx, int enable);
int  bmpc_ctx_profile_read(bmpc_ctx* ctx, int which, double* ms_total, uint64_t* launches);

/* ---- bases: SourceBuilder for (Arc<Vec<G>>, usize)  (src/multiexp.rs:45-86) -------------- */
/* Upload a base vector once; the (handle, offset) pair passed to bmpc_multiexp is the
 * reference's `(Arc<Vec<G::Affine>>, usize)` source (src/groth16/mod.rs:438-477). */
int  bmpc_bases_register(bmpc_ctx* ctx, int group, const void* points, size_t n, size_t stride,
                         int form, bmpc_bases** out);
/* same, from a device array of Montgomery x|y points (copied) */
int  bmpc_bases_register_dev(bmpc_ctx* ctx, int group, const void* d_points_mont, size_t n,
                             bmpc_bases** out, void* stream);
/* Precompute window tables 2^(c w) * P_i for w = 1 .. W-1 next to the bases (W x the memory, laid
 * out for the 180 GB of HBM): every window of a later multiexp then feeds ONE bucket set and the
 * top-down doubling fold of multiexp.rs:244-249 disappears.  window_bits = 0 picks c from n.
 * Results are unchanged (window-independent, SURVEY 8a'/7); call once per CRS vector. */
int  bmpc_bases_precompute(bmpc_ctx* ctx, bmpc_bases* b, int window_bits);
size_t bmpc_bases_len(const bmpc_bases* b);
int  bmpc_bases_group(const bmpc_bases* b);
/* read points [start, start+count) back as uncompressed big-endian (tests / CRS export) */
int  bmpc_bases_read(bmpc_ctx* ctx, const bmpc_bases* b, size_t start, size_t count, uint8_t* out);
/* device pointer of the Montgomery x|y array (for harnesses that keep data resident) */
const void* bmpc_bases_dev_ptr(const bmpc_bases* b);
void bmpc_bases_free(bmpc_ctx* ctx, bmpc_bases* b);

/* ---- multiexp  (src/multiexp.rs:254-281 `multiexp`, :159-250 `multiexp_inner`) ---------- */
/* result = sum over dense positions k-th -> bases[base_offset + k] * scalars[position].
 * density_words == NULL  <=> FullDensity; otherwise density_len must equal n
 * (the reference's assert at :273-278 -> BMPC_ERR_LENGTH_MISMATCH).
 * Error semantics follow SURVEY 8(a'): UNEXPECTED_EOF / UNEXPECTED_IDENTITY exactly when
 * the reference's Source::{next,skip} would fail, same precedence (highest failing window of
 * the reference's own window size, first error in scan order).
 * out: 96 B (G1) / 192 B (G2) uncompressed affine of the sum (G::to_affine().to_uncompressed()). */
int  bmpc_multiexp(bmpc_ctx* ctx, const bmpc_bases* bases, size_t base_offset,
                   const uint64_t* scalars, size_t n,
                   const uint64_t* density_words, size_t density_len, uint8_t* out);
int  bmpc_multiexp_dev(bmpc_ctx* ctx, const bmpc_bases* bases, size_t base_offset,
                       const uint64_t* d_scalars, size_t n,
                       const uint64_t* d_density_words, size_t density_len,
                       uint8_t* out, void* stream);
/* Sharded form: same, but the result is left as this rank's partial sum in
 * Montgomery XYZZ limbs (G1 192 B / G2 384 B) plus a status word; partials from all ranks are
 * all-gathered by the host and folded with bmpc_sum_partials. */
int  bmpc_multiexp_partial_dev(bmpc_ctx* ctx, const bmpc_bases* bases, size_t base_offset,
                               const uint64_t* d_scalars, size_t n,
                               const uint64_t* d_density_words, size_t density_len,
                               void* d_partial_out, void* stream);
/* One shard of a multiexp whose exponent vector has n_total (>= n) entries in all: same as
 * bmpc_multiexp_partial_dev, but the error conditions come back as the raw flag word instead of a
 * status, because which error the reference reports depends on ALL shards (multiexp.rs:244-249:
 * the first error in scan order of the highest failing window; the window size follows n_total,
 * :267-271).  The caller ORs the flag words of all shards and asks bmpc_msm_flags_status.
 * Returns BMPC_OK unless the call itself failed (arguments, CUDA). */
enum { BMPC_MSM_FLAG_EOF = 1,        /* a dense position maps past the end of the bases          */
       BMPC_MSM_FLAG_IDENT_ANY = 2,  /* a consumed base (dense, scalar != 0) is the identity     */
       BMPC_MSM_FLAG_IDENT_TOP = 4   /* ... and its digit in the reference's top window is != 0  */ };
int  bmpc_multiexp_shard_dev(bmpc_ctx* ctx, const bmpc_bases* bases, size_t base_offset,
                             const uint64_t* d_scalars, size_t n,
                             const uint64_t* d_density_words, size_t density_len, size_t n_total,
                             void* d_partial_out, uint32_t* flags_out, void* stream);
/* Enqueue-only shard call: no synchronisation.  The shard's record = its XYZZ partial followed by its raw
 * flag word (bmpc_shard_record_bytes(group) bytes, 16-byte aligned) is left at d_record_out in stream
 * order; all-gather the records of all ranks on the same stream and hand them to bmpc_fold_shard_records
 * -- one host synchronisation per multiexp instead of two (a rank's step at 8 GPUs is ~11 ms, the
 * synchronisations and launch gaps of the two-call form cost 0.2 ms of it). */
int  bmpc_multiexp_shard_enqueue_dev(bmpc_ctx* ctx, const bmpc_bases* bases, size_t base_offset,
                                     const uint64_t* d_scalars, size_t n,
                                     const uint64_t* d_density_words, size_t density_len, size_t n_total,
                                     void* d_record_out, void* stream);
size_t bmpc_shard_record_bytes(int group);
/* sum of the records' partials -> out (96 / 192 B uncompressed affine), OR of their flag words ->
 * *flags_or_out (status = bmpc_msm_flags_status); records are `stride` bytes apart */
int  bmpc_fold_shard_records(bmpc_ctx* ctx, int group, const void* d_records, size_t count, size_t stride,
                             uint8_t* out, uint32_t* flags_or_out, void* stream);
/* status of a whole multiexp from the OR of its shards' flag words (SURVEY 8a'/5) */
int  bmpc_msm_flags_status(uint32_t flags_or);
int  bmpc_sum_partials(bmpc_ctx* ctx, int group, const void* d_partials, size_t count,
                       uint8_t* out, void* stream);
size_t bmpc_partial_bytes(int group);
/* geometry the library will use for an n-point multiexp over `bases` (for reporting work done):
 * window bits c, number of windows W (point additions per dense point), bucket sets H */
int  bmpc_msm_geometry(bmpc_ctx* ctx, const bmpc_bases* bases, size_t n, uint32_t* window_bits,
                       uint32_t* windows, uint32_t* bucket_sets);
/* which bucket-accumulation kernel that multiexp runs: info[0] = 2 rounds of pair additions
 * (msm_pairs.cuh) | 1 batched-affine tree (msm_affine.cuh) | 0 XYZZ chain; info[1] = slices per job
 * (pair rounds: number of rounds), info[2] = additions per inversion per thread, info[3] = thread
 * blocks, info[4] = threads per block, info[5] = max points per slice */
int  bmpc_msm_accumulate_info(bmpc_ctx* ctx, const bmpc_bases* bases, size_t n, uint32_t info[8]);

/* Asynchronous form -- the reference's Waiter (src/multicore.rs:33-118): `multiexp` returns at once
 * and prover.rs:233-307 keeps eight of them in flight before the first wait().  The call enqueues
 * the upload of the scalars and the whole multiexp on a lane of the context (its own stream, scratch
 * arena and staging; up to 16 lanes, created on first use) and returns; `scalars` / `density_words`
 * must stay valid until bmpc_waiter_wait.  bmpc_waiter_wait blocks, writes the 96 / 192 result bytes,
 * returns the multiexp's status (same semantics as bmpc_multiexp) and frees the waiter.  With two or
 * more in flight the upload of one overlaps the kernels of the other. */
typedef struct bmpc_waiter bmpc_waiter;
int  bmpc_multiexp_async(bmpc_ctx* ctx, const bmpc_bases* bases, size_t base_offset,
                         const uint64_t* scalars, size_t n,
                         const uint64_t* density_words, size_t density_len, bmpc_waiter** out);
int  bmpc_waiter_wait(bmpc_waiter* w, uint8_t* out);

/* ---- EvaluationDomain  (src/domain.rs:21-189) ------------------------------------------ */
/* from_coeffs (:47-79): pads with zeros to m = 2^exp >= len (m = 1 for len <= 1);
 * exp >= 32 -> BMPC_ERR_DEGREE_TOO_LARGE.  coeffs: len x 4 u64 Montgomery limbs (host). */
int  bmpc_domain_from_coeffs(bmpc_ctx* ctx, const uint64_t* coeffs, size_t len, bmpc_domain** out);
int  bmpc_domain_from_coeffs_dev(bmpc_ctx* ctx, const uint64_t* d_coeffs, size_t len,
                                 bmpc_domain** out, void* stream);
size_t   bmpc_domain_len(const bmpc_domain* d);      /* m          */
uint32_t bmpc_domain_exp(const bmpc_domain* d);      /* log2 m     */
/* into_coeffs (:43-45): copy the m coefficients back to the host */
int  bmpc_domain_into_coeffs(bmpc_ctx* ctx, const bmpc_domain* d, uint64_t* out);
uint64_t* bmpc_domain_dev_ptr(bmpc_domain* d);
void bmpc_domain_free(bmpc_ctx* ctx, bmpc_domain* d);
/* fft / ifft / coset_fft / icoset_fft (:81-125); op = BMPC_FFT ... */
int  bmpc_domain_transform(bmpc_ctx* ctx, bmpc_domain* d, int op, void* stream);
/* distribute_powers(g) (:101-113): coeffs[i] *= g^i ; g = 4 x u64 Montgomery */
int  bmpc_domain_distribute_powers(bmpc_ctx* ctx, bmpc_domain* d, const uint64_t g[4], void* stream);
/* z(tau) = tau^m - 1 (:129-134) ; tau/out Montgomery */
int  bmpc_domain_z(bmpc_ctx* ctx, const bmpc_domain* d, const uint64_t tau[4], uint64_t out[4]);
/* divide_by_z_on_coset (:139-151) */
int  bmpc_domain_divide_by_z_on_coset(bmpc_ctx* ctx, bmpc_domain* d, void* stream);
/* mul_assign / sub_assign (:154-189); length mismatch -> BMPC_ERR_LENGTH_MISMATCH */
int  bmpc_domain_mul_assign(bmpc_ctx* ctx, bmpc_domain* d, const bmpc_domain* other, void* stream);
int  bmpc_domain_sub_assign(bmpc_ctx* ctx, bmpc_domain* d, const bmpc_domain* other, void* stream);
/* In-place transform of a caller-owned device array of m = 2^log_m Montgomery coefficients
 * (harness entry point: inputs resident in HBM). */
int  bmpc_ntt_dev(bmpc_ctx* ctx, uint64_t* d_coeffs, uint32_t log_m, int op, void* stream);
/* Host-buffer convenience: upload, transform, download (what a drop-in `fft(&worker)` does). */
int  bmpc_ntt(bmpc_ctx* ctx, uint64_t* coeffs, uint32_t log_m, int op);

/* ---- pieces of the distributed (multi-GPU) transform ----------------------------------------
 * A domain whose coefficients are split over G GPUs (rank g holds the contiguous slice
 * [g m/G, (g+1) m/G)) is transformed as a four-step decomposition m = R x C: transpose, R-point
 * column transforms, twiddle by omega_m^(j2 k1), transpose, C-point row transforms, transpose back
 * to natural order -- the same outputs as best_fft (src/domain.rs:261-372) on the whole vector.
 * The exchange between the steps is an all-to-all over NCCL, driven by the host layer
 * (bellman_mpc_b200/dist.py: DistributedDomain); these entry points are the per-GPU steps. */
/* `batch` (<= 65535) independent plain transforms of 2^log_n coefficients, laid out back to back,
 * in place; inverse != 0 uses omega^-1 and does NOT scale by 1/n. */
int  bmpc_ntt_batch_dev(bmpc_ctx* ctx, uint64_t* d_coeffs, uint32_t log_n, uint32_t batch, int inverse, void* stream);
/* out[b][a][c] = in[a][b][c] over 32-byte coefficients, a < d0, b < d1, c < d2 (d2 == 1: a plain
 * transpose); in and out must not overlap. */
int  bmpc_fr_swap01_dev(bmpc_ctx* ctx, const uint64_t* d_in, uint64_t* d_out, uint32_t d0, uint32_t d1, uint32_t d2,
                        void* stream);
/* d[r][c] *= w^((row0 + r) c), r < rows, c < cols, w = omega_m (inverse != 0: omega_m^-1), m = 2^log_m */
int  bmpc_ntt_fourstep_twiddle_dev(bmpc_ctx* ctx, uint64_t* d, uint32_t rows, uint32_t cols, uint32_t row0,
                                   uint32_t log_m, int inverse, void* stream);
/* d[i] *= f(first + i), i < n, for the domain of size 2^log_m: which = 0: g^i (distribute_powers
 * of coset_fft, :116-119), 1: g^-i / m (icoset_fft, :121-125), 2: 1 / m (ifft, :88-98) */
int  bmpc_fr_scale_pow_dev(bmpc_ctx* ctx, uint64_t* d, size_t n, uint32_t first, uint32_t log_m, int which,
                           void* stream);

/* ---- H polynomial  (src/groth16/prover.rs:210-231) -------------------------------------- */
/* a, b, c: evaluations (len x 4 u64 Montgomery, host).  Runs 3 x (ifft, coset_fft),
 * a*b - c, divide_by_z_on_coset, icoset_fft, drops the last coefficient and converts to
 * canonical form (`to_le_bits`), fused on the device.  out: (m - 1) x 4 u64 canonical. */
int  bmpc_h_coefficients(bmpc_ctx* ctx, const uint64_t* a, const uint64_t* b, const uint64_t* c,
                         size_t len, uint64_t* out, size_t* out_len);
/* device-resident form: d_a/d_b/d_c each hold m = 2^log_m coefficients (already padded) and
 * are clobbered; the m-1 canonical scalars are left in d_a. */
int  bmpc_h_coefficients_dev(bmpc_ctx* ctx, uint64_t* d_a, uint64_t* d_b, uint64_t* d_c,
                             uint32_t log_m, void* stream);
/* The same pipeline in its two halves, so that several GPUs can share it (one vector each, one device
 * combines; bellman_mpc_b200/dist.py):
 *   bmpc_h_coset_evals_dev: d_p (m Montgomery evaluations, padded) <- ifft, coset_fft (prover.rs:214-219
 *     for one of a, b, c): the polynomial's evaluations on the coset; host_src != NULL: the host_len
 *     evaluations are first uploaded from there and padded with zeros to m (from_coeffs, :211-213);
 *   bmpc_h_from_coset_evals_dev: d_a <- (d_a * d_b - d_c) / Z on the coset, icoset_fft, to_le_bits
 *     (prover.rs:221-231): the m-1 canonical H scalars are left in d_a (entry m-1 is dropped by the caller).
 * bmpc_h_coefficients_dev == three times the first, then the second. */
int  bmpc_h_coset_evals_dev(bmpc_ctx* ctx, uint64_t* d_p, uint32_t log_m, const uint64_t* host_src,
                            size_t host_len, void* stream);
int  bmpc_h_from_coset_evals_dev(bmpc_ctx* ctx, uint64_t* d_a, const uint64_t* d_b, const uint64_t* d_c,
                                 uint32_t log_m, void* stream);
/* Montgomery -> canonical (PrimeFieldBits::to_le_bits, prover.rs:231,241,248), in place */
int  bmpc_fr_to_canonical_dev(bmpc_ctx* ctx, uint64_t* d_vals, size_t n, void* stream);

/* ---- create_proof  (src/groth16/prover.rs:176-350, everything after synthesis) ---------- */
typedef struct {
    /* ParameterSource for &Parameters (src/groth16/mod.rs:438-477) */
    const bmpc_bases* h;
    const bmpc_bases* l;
    const bmpc_bases* a;
    const bmpc_bases* b_g1;
    const bmpc_bases* b_g2;
    /* VerifyingKey parts the prover touches (prover.rs:309-327), uncompressed big-endian */
    uint8_t alpha_g1[96];
    uint8_t beta_g1[96];
    uint8_t beta_g2[192];
    uint8_t delta_g1[96];
    uint8_t delta_g2[192];
} bmpc_params;

typedef struct {
    /* ProvingAssignment after synthesis (prover.rs:55-69); Fr values are Montgomery limbs */
    const uint64_t* a;  const uint64_t* b;  const uint64_t* c;   /* num_constraints x 4 */
    size_t num_constraints;
    const uint64_t* input_assignment;  size_t num_inputs;        /* x 4 */
    const uint64_t* aux_assignment;    size_t num_aux;           /* x 4 */
    const uint64_t* a_aux_density;     /* num_aux bits   */
    const uint64_t* b_input_density;   /* num_inputs bits */
    const uint64_t* b_aux_density;     /* num_aux bits   */
} bmpc_assignment;

/* r, s: 4 x u64 Montgomery.  proof_out: 192 B = A (G1 compressed) | B (G2 compressed) | C
 * (Proof::write, src/groth16/mod.rs:42-48). */
int  bmpc_create_proof(bmpc_ctx* ctx, const bmpc_params* params, const bmpc_assignment* asg,
                       const uint64_t r[4], const uint64_t s[4], uint8_t proof_out[192]);

/* create_proof sharded over the GPUs of a node: every multiexp's exponent range is split
 * into one contiguous slice per rank; the rank holds the matching slice of each query vector.
 * Order of the eight multiexps everywhere below: a_inputs, a_aux, b_g1_inputs, b_g1_aux,
 * b_g2_inputs, b_g2_aux, h, l (the order prover.rs:328-343 awaits them, h and l last).
 *   bmpc_create_proof_partials: `asg` carries the rank's SLICE of the assignment (input/aux
 *     pointers, counts and density words already offset; a, b, c complete: the H polynomial of
 *     prover.rs:210-231 is recomputed on every rank, which costs no collective and no more wall
 *     time than computing it once and scattering it); `params` the rank's slices of the query
 *     vectors; shard->base_offset[j] = the first base multiexp j consumes inside that slice;
 *     [h_lo, h_hi) = the rank's share of the m-1 H coefficients.  Returns the eight XYZZ partial
 *     sums (6 x 192 B G1, then 2 x 384 B G2) and the eight multiexp statuses.
 *   bmpc_create_proof_finish: `partials_all` = world x BMPC_PROOF_PARTIAL_BYTES gathered from all
 *     ranks (any collective: 1920 B per rank); folds them, runs prover.rs:309-349, writes the proof.
 *     The caller first ORs the flag words per multiexp over the ranks and resolves them with
 *     bmpc_msm_flags_status in the order above (bellman_mpc_b200/dist.py::combine_flags). */
#define BMPC_PROOF_PARTIAL_BYTES 1920
typedef struct {
    size_t base_offset[8];
    size_t h_lo, h_hi;
    size_t n_total[8];
} bmpc_proof_shard;
int  bmpc_create_proof_partials(bmpc_ctx* ctx, const bmpc_params* params, const bmpc_assignment* asg,
                                const bmpc_proof_shard* shard, uint8_t partials_out[BMPC_PROOF_PARTIAL_BYTES],
                                uint32_t flags_out[8]);
int  bmpc_create_proof_finish(bmpc_ctx* ctx, const bmpc_params* params, const uint8_t* partials_all, size_t world,
                              const uint64_t r[4], const uint64_t s[4], uint8_t proof_out[192]);

/* ---- all GPUs of a node behind one call (SURVEY 8b `bmpc_ctx_create(devices, n)`, 8e) ---------------
 * The reference's multiexp (multiexp.rs:254-281) and create_proof (prover.rs:176-350) are one call
 * in one process; bmpc_multi keeps that shape over N devices: one context per device inside, a base
 * vector split evenly by base index at registration (device g holds bases [g n/N, (g+1) n/N)), and
 * per call the exponent positions cut so that each device's positions consume exactly its bases
 * (with a density map: the position of the k-th dense bit, counted on the host), one host thread
 * per device, the XYZZ partial sums gathered on device 0 by peer copies (cudaMemcpyPeerAsync over
 * NVLink, 192 / 384 bytes per device) and folded there, and the raw error flags of all shards ORed
 * before the reference's precedence rule is applied (multiexp.rs:244-249).  Same result bytes and
 * statuses as the single-device calls.  `devices` may name a device more than once (tests). */
typedef struct bmpc_multi bmpc_multi;
typedef struct bmpc_multi_bases bmpc_multi_bases;
int  bmpc_multi_create(const int* devices, int n, bmpc_multi** out);
void bmpc_multi_destroy(bmpc_multi* m);
int  bmpc_multi_size(const bmpc_multi* m);
bmpc_ctx* bmpc_multi_ctx(bmpc_multi* m, int rank);      /* the per-device context (tuning, profiling) */
const char* bmpc_multi_last_error(const bmpc_multi* m);
int  bmpc_multi_bases_register(bmpc_multi* m, int group, const void* points, size_t n, size_t stride,
                               int form, bmpc_multi_bases** out);
int  bmpc_multi_bases_precompute(bmpc_multi* m, bmpc_multi_bases* b, int window_bits);
size_t bmpc_multi_bases_len(const bmpc_multi_bases* b);
/* device `rank`'s slice and the index of its first base in the whole vector */
const bmpc_bases* bmpc_multi_bases_part(const bmpc_multi_bases* b, int rank, size_t* first);
void bmpc_multi_bases_free(bmpc_multi* m, bmpc_multi_bases* b);
/* multiexp.rs:254-281 over all devices; arguments and statuses as bmpc_multiexp (host pointers) */
int  bmpc_multi_multiexp(bmpc_multi* m, const bmpc_multi_bases* bases, size_t base_offset,
                         const uint64_t* scalars, size_t n,
                         const uint64_t* density_words, size_t density_len, uint8_t* out);
typedef struct {
    const bmpc_multi_bases* h;
    const bmpc_multi_bases* l;
    const bmpc_multi_bases* a;
    const bmpc_multi_bases* b_g1;
    const bmpc_multi_bases* b_g2;
    uint8_t alpha_g1[96];
    uint8_t beta_g1[96];
    uint8_t beta_g2[192];
    uint8_t delta_g1[96];
    uint8_t delta_g2[192];
} bmpc_multi_params;
/* prover.rs:206-350 over all devices: every device runs its share of the eight multiexps; the H
 * polynomial is computed once between them (devices 0-2 upload and transform one of a, b, c each,
 * device 0 pulls the other two over NVLink, finishes the pipeline, every device pulls its slice of the
 * scalars; second contexts, under the other seven multiexps: 8 GPUs, 2^22: 0.051 -> 0.034 s); the 1920
 * bytes of partial sums per device are folded and the tail runs on device 0.  Same 192 proof bytes
 * and statuses as bmpc_create_proof. */
int  bmpc_multi_create_proof(bmpc_multi* m, const bmpc_multi_params* params, const bmpc_assignment* asg,
                             const uint64_t r[4], const uint64_t s[4], uint8_t proof_out[192]);

/* ---- Parameters / VerifyingKey wire format  (src/groth16/mod.rs:146-221,261-400) ------------- */
typedef struct {
    bmpc_params p;            /* query vectors resident in HBM + the vk points the prover uses    */
    uint8_t gamma_g2[192];    /* the rest of the VerifyingKey                                     */
    bmpc_bases* ic;
} bmpc_parameters;
/* Parameters::read(reader, checked): parses the blob, decodes and validates every point on the
 * device (flags, canonical coordinates; vk and ic always, the query vectors when `checked`: on the
 * curve and in the prime-order subgroup; identity rejected in ic and the query vectors), leaves
 * the vectors resident.  Errors in stream order like the reference: BMPC_ERR_INVALID_DATA
 * (bmpc_last_error = "invalid G1" | "invalid G2" | "point at infinity") or
 * BMPC_ERR_UNEXPECTED_EOF for a short blob. */
int  bmpc_params_read(bmpc_ctx* ctx, const uint8_t* data, size_t len, int checked, bmpc_parameters* out);
/* Parameters::write: *written <- size; BMPC_ERR_INVALID with *written set if cap is too small */
int  bmpc_params_write(bmpc_ctx* ctx, const bmpc_parameters* in, uint8_t* out, size_t cap, size_t* written);
void bmpc_params_free(bmpc_ctx* ctx, bmpc_parameters* p);

/* ---- constraint-system side -------------------------------------------------- */
/* R1CS matrix in CSR form.  For bmpc_r1cs_eval: one row per constraint, columns = variables
 * ([0, num_inputs) inputs with column 0 = ONE, then aux).  For bmpc_generate_parameters: the
 * TRANSPOSE -- one row per variable listing (coeff, constraint), the KeypairAssembly layout of
 * generator.rs:44-156.  coeff: Montgomery 4 x u64. */
typedef struct {
    const uint32_t* row_ptr;   /* num_rows + 1 */
    const uint32_t* col;       /* nnz */
    const uint64_t* coeff;     /* nnz x 4 */
    size_t num_rows, nnz;
} bmpc_csr;
/* ProvingAssignment::enforce for the whole system (prover.rs:19-53,100-138) plus the `x * 0 = 0`
 * rows create_proof appends per input (:202-204).  Outputs (host): a, b, c of
 * (A->num_rows + num_inputs) x 4 u64 Montgomery, and the three density maps (bitvec words). */
int  bmpc_r1cs_eval(bmpc_ctx* ctx, const bmpc_csr* A, const bmpc_csr* B, const bmpc_csr* C, size_t num_inputs,
                    size_t num_aux, const uint64_t* input_assignment, const uint64_t* aux_assignment,
                    uint64_t* a_out, uint64_t* b_out, uint64_t* c_out, uint64_t* a_aux_density,
                    uint64_t* b_input_density, uint64_t* b_aux_density);
/* generate_parameters with upstream semantics (generator.rs:241-634 minus the fork's MPC
 * cross-check hooks): num_constraints = user constraints (the input rows are appended here, :273-275);
 * g1/g2 uncompressed generators; scalars Montgomery.  BMPC_ERR_UNEXPECTED_IDENTITY if gamma or delta
 * is zero (:330-345); BMPC_ERR_INVALID_DATA ("UnconstrainedVariable", :584-590). */
int  bmpc_generate_parameters(bmpc_ctx* ctx, const bmpc_csr* At, const bmpc_csr* Bt, const bmpc_csr* Ct,
                              size_t num_inputs, size_t num_aux, size_t num_constraints, const uint8_t g1[96],
                              const uint8_t g2[192], const uint64_t alpha[4], const uint64_t beta[4],
                              const uint64_t gamma[4], const uint64_t delta[4], const uint64_t tau[4],
                              bmpc_parameters* out);

/* ---- batch scalar multiplication  (src/groth16/mpc.rs:647-706 make_new_paramter /
 *      make_new_tau_paramter; also fixed-base CRS synthesis, generator.rs:372-397,492-512) -- */
/* out[i] = in[i] * k[i] (per_element = 1) or in[i] * k[0] (per_element = 0); k canonical 4 x u64 */
int  bmpc_batch_scalar_mul(bmpc_ctx* ctx, const bmpc_bases* in, const uint64_t* scalars,
                           int per_element, bmpc_bases** out);
/* list_mul_matrix (src/groth16/mpc.rs:416-457), one group at a time (the reference runs the same
 * loop over a G1 and a G2 list): out has list.len() elements, out[i] = sum_j coeffs[j] * list[cols[j]]
 * over CSR row i (entries row_ptr[i] .. row_ptr[i+1]) for the rows BEFORE the first empty one (the
 * reference `break`s there, :432-434); every later element is the identity.  coeffs canonical 4 x u64
 * per entry.  BMPC_ERR_LENGTH_MISMATCH where the reference panics on an index: more live rows than
 * list.len() (a taller matrix whose first empty row comes before row list.len() is fine, as in the
 * reference) or a live column >= list.len(). */
int  bmpc_list_mul_matrix(bmpc_ctx* ctx, const bmpc_bases* list, const uint64_t* row_ptr,
                          const uint32_t* cols, const uint64_t* coeffs, size_t n_rows, bmpc_bases** out);
/* out[i] = base * k[i]; base uncompressed big-endian; scalars canonical (host or device) */
int  bmpc_fixed_base_mul(bmpc_ctx* ctx, int group, const uint8_t* base, const uint64_t* scalars,
                         size_t n, int scalars_on_device, bmpc_bases** out);

#ifdef __cplusplus
}
#endif
#endif /* BELLMAN_B200_H */
