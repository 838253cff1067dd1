/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference's hot path.
 *
 * Restates, in plain C with pthreads, the algorithms of /root/reference/bellman/src:
 *   multiexp.rs:159-281   multiexp / multiexp_inner: window c = 3 if n < 32 else ceil(ln n),
 *                          one task per window on the pool, 2^c - 1 buckets, summation by
 *                          parts, top-down Horner fold, density / offset / skip semantics
 *   domain.rs:47-189       EvaluationDomain ops;  :261-372 best_fft / serial_fft / parallel_fft
 *   multicore.rs:29-31,78-91,120-130  log_num_threads, chunking of Worker::scope
 *   groth16/prover.rs:206-350  create_proof after synthesis (7 transforms, 8 multiexps in
 *                          flight, tail algebra)
 * over a 64-bit-limb Montgomery restatement of bls12_381 0.6.0 (field.h, curve_tmpl.h).
 *
 * It is the parity checker for sizes the Python oracle cannot reach and the timed "CPU
 * restatement of reference" baseline (the Rust crate cannot be built here: no cargo, its
 * arithmetic crates are not vendored).  Pinned against oracle/*.py (itself pinned to the
 * reference's dummy-engine golden vectors) by tests/test_oracle_c.py.  Never linked into or
 * loaded by the product library.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "field.h"

/* ------------------------------------------------------------------ group instantiations */
static inline void fp_mul_3b(fp_t* r, const fp_t* a) { /* 3b = 12 on G1 */
    fp_t t2, t4, t8;
    fp_dbl(&t2, a); fp_dbl(&t4, &t2); fp_dbl(&t8, &t4);
    fp_add(r, &t8, &t4);
}
static inline void fp2_mul_3b(fp2_t* r, const fp2_t* a) { /* 3b = 12 (1 + u) on G2 */
    fp_t d, s;
    fp_sub(&d, &a->c0, &a->c1);
    fp_add(&s, &a->c0, &a->c1);
    fp_mul_3b(&r->c0, &d);
    fp_mul_3b(&r->c1, &s);
}

#define FE fp_t
#define FE_(op) fp_##op
#define PT g1_t
#define AF g1_affine_t
#define GP_(f) g1_##f
#define MUL_3B fp_mul_3b
#include "curve_tmpl.h"
#undef FE
#undef FE_
#undef PT
#undef AF
#undef GP_
#undef MUL_3B

#define FE fp2_t
#define FE_(op) fp2_##op
#define PT g2_t
#define AF g2_affine_t
#define GP_(f) g2_##f
#define MUL_3B fp2_mul_3b
#include "curve_tmpl.h"
#undef FE
#undef FE_
#undef PT
#undef AF
#undef GP_
#undef MUL_3B

/* G1 generator, Montgomery limbs (y checked by tests/test_oracle_c.py) */
static const uint64_t G1_GEN_X[6] = {0x5cb38790fd530c16ULL, 0x7817fc679976fff5ULL, 0x154f95c7143ba1c1ULL,
                                     0xf0ae6acdf3d0e747ULL, 0xedce6ecc21dbf440ULL, 0x120177419e0bfb75ULL};
static const uint64_t G1_GEN_Y[6] = {0xbaac93d50ce72271ULL, 0x8c22631a7918fd8eULL, 0xdd595f13570725ceULL,
                                     0x51ac582950405194ULL, 0x0e1c8c3fad0059c0ULL, 0x0bbc3efc5008a26aULL};

/* ------------------------------------------------------------------------- tiny task pool
 * rayon stand-in: run `count` independent tasks on up to `threads` OS threads. */
typedef void (*task_fn)(void* arg, size_t idx);
typedef struct { task_fn fn; void* arg; size_t count; size_t next; pthread_mutex_t mu; } job_t;

static void* job_worker(void* p) {
    job_t* j = (job_t*)p;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        size_t i = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (i >= j->count) break;
        j->fn(j->arg, i);
    }
    return NULL;
}
static void run_tasks(task_fn fn, void* arg, size_t count, int threads) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > count) threads = (int)count;
    job_t j = {fn, arg, count, 0, PTHREAD_MUTEX_INITIALIZER};
    if (threads <= 1) { job_worker(&j); return; }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, job_worker, &j);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
}

static uint32_t log2_floor(size_t num) { /* multicore.rs:120-130 */
    uint32_t pow = 0;
    while (((size_t)1 << (pow + 1)) <= num) pow++;
    return pow;
}

/* ------------------------------------------------------------------------------ multiexp */
uint32_t orc_window_size(size_t n) { /* multiexp.rs:267-271 */
    if (n < 32) return 3;
    return (uint32_t)ceil(log((double)(uint32_t)n));
}

typedef struct {
    int group; /* 1 = G1, 2 = G2 */
    const void* bases; size_t nbases, start;
    const uint64_t* exps; size_t n;
    const uint64_t* density;
    uint32_t c;
    void* parts; int* errs;
} mexp_job_t;

static void mexp_task(void* arg, size_t idx) {
    mexp_job_t* j = (mexp_job_t*)arg;
    uint32_t skip = (uint32_t)idx * j->c;
    if (j->group == 1)
        j->errs[idx] = g1_multiexp_region((const g1_affine_t*)j->bases, j->nbases, j->start, j->exps, j->n,
                                          j->density, skip, j->c, (g1_t*)j->parts + idx);
    else
        j->errs[idx] = g2_multiexp_region((const g2_affine_t*)j->bases, j->nbases, j->start, j->exps, j->n,
                                          j->density, skip, j->c, (g2_t*)j->parts + idx);
}

/* multiexp_inner + the fold (multiexp.rs:238-249).  out: g1_t or g2_t.  Returns status. */
static int multiexp_any_c(int group, const void* bases, size_t nbases, size_t start, const uint64_t* exps,
                          size_t n, const uint64_t* density, int threads, uint32_t c, void* out);
static int multiexp_any(int group, const void* bases, size_t nbases, size_t start, const uint64_t* exps,
                        size_t n, const uint64_t* density, int threads, void* out) {
    return multiexp_any_c(group, bases, nbases, start, exps, n, density, threads, orc_window_size(n), out);
}
/* c given by the caller: bench.py times a SAMPLE of a larger exponent vector with the window the
 * reference picks for the whole vector (multiexp.rs:267-271), so the per-point work is the stated
 * configuration's. */
static int multiexp_any_c(int group, const void* bases, size_t nbases, size_t start, const uint64_t* exps,
                          size_t n, const uint64_t* density, int threads, uint32_t c, void* out) {
    size_t nwin = (255 + c - 1) / c; /* (0..NUM_BITS).step_by(c) */
    size_t psz = group == 1 ? sizeof(g1_t) : sizeof(g2_t);
    void* parts = malloc(nwin * psz);
    int* errs = (int*)calloc(nwin, sizeof(int));
    mexp_job_t j = {group, bases, nbases, start, exps, n, density, c, parts, errs};
    run_tasks(mexp_task, &j, nwin, threads);
    int status = 0;
    if (group == 1) {
        g1_t acc; g1_identity(&acc);
        for (size_t w = nwin; w-- > 0;) {
            if (errs[w]) { status = errs[w]; break; }
            for (uint32_t k = 0; k < c; k++) g1_dbl(&acc, &acc);
            g1_add(&acc, &acc, (g1_t*)parts + w);
        }
        if (!status) *(g1_t*)out = acc;
    } else {
        g2_t acc; g2_identity(&acc);
        for (size_t w = nwin; w-- > 0;) {
            if (errs[w]) { status = errs[w]; break; }
            for (uint32_t k = 0; k < c; k++) g2_dbl(&acc, &acc);
            g2_add(&acc, &acc, (g2_t*)parts + w);
        }
        if (!status) *(g2_t*)out = acc;
    }
    free(parts); free(errs);
    return status;
}

/* ---- encodings ------------------------------------------------------------------------
 * "mont xy" = raw Montgomery limbs x|y, all-zero = identity (the GPU library's resident form);
 * uncompressed = ZCash big-endian (G1 96 B, G2 192 B c1|c0). */
static void g1_affine_from_mont(g1_affine_t* a, const uint64_t* p) {
    memcpy(a->x.l, p, 48); memcpy(a->y.l, p + 6, 48);
    a->inf = fp_is_zero(&a->x) && fp_is_zero(&a->y);
}
static void g2_affine_from_mont(g2_affine_t* a, const uint64_t* p) {
    memcpy(&a->x, p, 96); memcpy(&a->y, p + 12, 96);
    a->inf = fp2_is_zero(&a->x) && fp2_is_zero(&a->y);
}
static void fp_to_be(const fp_t* m, uint8_t* out) {
    fp_t c; fp_from_mont(&c, m);
    for (int i = 0; i < 6; i++)
        for (int b = 0; b < 8; b++) out[8 * i + b] = (uint8_t)(c.l[5 - i] >> (56 - 8 * b));
}
static void fp_from_be(fp_t* m, const uint8_t* in, uint8_t mask0) {
    fp_t c;
    for (int i = 0; i < 6; i++) {
        uint64_t v = 0;
        for (int b = 0; b < 8; b++) {
            uint8_t byte = in[8 * i + b];
            if (i == 0 && b == 0) byte &= mask0;
            v = (v << 8) | byte;
        }
        c.l[5 - i] = v;
    }
    fp_to_mont(m, &c);
}
void orc_g1_to_uncompressed(const g1_affine_t* a, uint8_t out[96]) {
    if (a->inf) { memset(out, 0, 96); out[0] = 0x40; return; }
    fp_to_be(&a->x, out); fp_to_be(&a->y, out + 48);
}
void orc_g2_to_uncompressed(const g2_affine_t* a, uint8_t out[192]) {
    if (a->inf) { memset(out, 0, 192); out[0] = 0x40; return; }
    fp_to_be(&a->x.c1, out); fp_to_be(&a->x.c0, out + 48);
    fp_to_be(&a->y.c1, out + 96); fp_to_be(&a->y.c0, out + 144);
}
static void g1_from_uncompressed(g1_affine_t* a, const uint8_t* in) {
    if (in[0] & 0x40) { fp_zero(&a->x); fp_zero(&a->y); a->inf = 1; return; }
    fp_from_be(&a->x, in, 0x1f); fp_from_be(&a->y, in + 48, 0xff); a->inf = 0;
}
static void g2_from_uncompressed(g2_affine_t* a, const uint8_t* in) {
    if (in[0] & 0x40) { fp2_zero(&a->x); fp2_zero(&a->y); a->inf = 1; return; }
    fp_from_be(&a->x.c1, in, 0x1f); fp_from_be(&a->x.c0, in + 48, 0xff);
    fp_from_be(&a->y.c1, in + 96, 0xff); fp_from_be(&a->y.c0, in + 144, 0xff); a->inf = 0;
}

/* Opaque base vectors so that repeated baseline runs do not re-decode. */
typedef struct { int group; size_t n; void* pts; } orc_bases_t;

orc_bases_t* orc_bases_from_uncompressed(int group, const uint8_t* data, size_t n) {
    orc_bases_t* b = (orc_bases_t*)malloc(sizeof(*b));
    b->group = group; b->n = n;
    if (group == 1) {
        g1_affine_t* p = (g1_affine_t*)malloc((n ? n : 1) * sizeof(g1_affine_t));
        for (size_t i = 0; i < n; i++) g1_from_uncompressed(&p[i], data + 96 * i);
        b->pts = p;
    } else {
        g2_affine_t* p = (g2_affine_t*)malloc((n ? n : 1) * sizeof(g2_affine_t));
        for (size_t i = 0; i < n; i++) g2_from_uncompressed(&p[i], data + 192 * i);
        b->pts = p;
    }
    return b;
}
orc_bases_t* orc_bases_from_mont(int group, const uint64_t* data, size_t n) {
    orc_bases_t* b = (orc_bases_t*)malloc(sizeof(*b));
    b->group = group; b->n = n;
    if (group == 1) {
        g1_affine_t* p = (g1_affine_t*)malloc((n ? n : 1) * sizeof(g1_affine_t));
        for (size_t i = 0; i < n; i++) g1_affine_from_mont(&p[i], data + 12 * i);
        b->pts = p;
    } else {
        g2_affine_t* p = (g2_affine_t*)malloc((n ? n : 1) * sizeof(g2_affine_t));
        for (size_t i = 0; i < n; i++) g2_affine_from_mont(&p[i], data + 24 * i);
        b->pts = p;
    }
    return b;
}
/* Synthetic G1 bases with known discrete logs for the CPU-only reference arm of bench.py:
 * P_i = (first + i) * G, built by repeated mixed addition and one batch inversion. */
static orc_bases_t* bases_g1_progression(size_t n, const uint64_t first[4], const uint64_t step[4]);
orc_bases_t* orc_bases_g1_sequence(size_t n, uint64_t first) {
    uint64_t a[4] = {first, 0, 0, 0}, one[4] = {1, 0, 0, 0};
    return bases_g1_progression(n, a, one);
}
/* P_i = (first + i * step) * G for 256-bit first, step: bases whose discrete logs are spread over
 * the whole scalar field (the reference arm's "uniform random bases"), still known in closed form */
orc_bases_t* orc_bases_g1_progression(size_t n, const uint64_t first[4], const uint64_t step[4]) {
    return bases_g1_progression(n, first, step);
}
static orc_bases_t* bases_g1_progression(size_t n, const uint64_t first[4], const uint64_t step[4]) {
    orc_bases_t* b = (orc_bases_t*)malloc(sizeof(*b));
    b->group = 1; b->n = n;
    g1_affine_t* out = (g1_affine_t*)malloc((n ? n : 1) * sizeof(g1_affine_t));
    b->pts = out;
    if (!n) return b;
    g1_affine_t gen; memcpy(gen.x.l, G1_GEN_X, 48); memcpy(gen.y.l, G1_GEN_Y, 48); gen.inf = 0;
    g1_t g, cur, stp; g1_from_affine(&g, &gen);
    g1_mul(&cur, &g, first);
    g1_mul(&stp, &g, step);
    g1_affine_t stp_a; g1_to_affine(&stp_a, &stp);
    gen = stp_a;                      /* the loop below adds `gen` per element */
    g1_t* proj = (g1_t*)malloc(n * sizeof(g1_t));
    fp_t* pre = (fp_t*)malloc(n * sizeof(fp_t));
    fp_t acc; fp_one(&acc);
    for (size_t i = 0; i < n; i++) {
        proj[i] = cur;
        pre[i] = acc;
        if (!g1_is_identity(&cur)) fp_mul(&acc, &acc, &cur.z);
        g1_add_mixed(&cur, &cur, &gen);
    }
    fp_t inv; fp_inv(&inv, &acc);
    for (size_t i = n; i-- > 0;) {
        if (g1_is_identity(&proj[i])) { fp_zero(&out[i].x); fp_zero(&out[i].y); out[i].inf = 1; continue; }
        fp_t zi; fp_mul(&zi, &inv, &pre[i]);
        fp_mul(&inv, &inv, &proj[i].z);
        fp_mul(&out[i].x, &proj[i].x, &zi);
        fp_mul(&out[i].y, &proj[i].y, &zi);
        out[i].inf = 0;
    }
    free(proj); free(pre);
    return b;
}
void orc_bases_free(orc_bases_t* b) { if (b) { free(b->pts); free(b); } }
size_t orc_bases_len(const orc_bases_t* b) { return b->n; }

/* multiexp(pool, (bases, start), density, exponents) -> uncompressed affine; status as
 * include/bellman_b200.h (0 ok, 1 UnexpectedIdentity, 2 UnexpectedEof, 4 length mismatch). */
int orc_multiexp(const orc_bases_t* b, size_t start, const uint64_t* exps, size_t n,
                 const uint64_t* density, size_t density_len, int threads, uint8_t* out) {
    if (density && density_len != n) return 4; /* multiexp.rs:273-278 */
    if (b->group == 1) {
        g1_t r; g1_affine_t a;
        int st = multiexp_any(1, b->pts, b->n, start, exps, n, density, threads, &r);
        if (st) return st;
        g1_to_affine(&a, &r); orc_g1_to_uncompressed(&a, out);
    } else {
        g2_t r; g2_affine_t a;
        int st = multiexp_any(2, b->pts, b->n, start, exps, n, density, threads, &r);
        if (st) return st;
        g2_to_affine(&a, &r); orc_g2_to_uncompressed(&a, out);
    }
    return 0;
}

/* same with the window of an n_window-entry exponent vector (n_window >= n; G1, FullDensity use) */
int orc_multiexp_window(const orc_bases_t* b, size_t start, const uint64_t* exps, size_t n,
                        size_t n_window, int threads, uint8_t* out) {
    uint32_t c = orc_window_size(n_window);
    if (b->group == 1) {
        g1_t r; g1_affine_t a;
        int st = multiexp_any_c(1, b->pts, b->n, start, exps, n, NULL, threads, c, &r);
        if (st) return st;
        g1_to_affine(&a, &r); orc_g1_to_uncompressed(&a, out);
    } else {
        g2_t r; g2_affine_t a;
        int st = multiexp_any_c(2, b->pts, b->n, start, exps, n, NULL, threads, c, &r);
        if (st) return st;
        g2_to_affine(&a, &r); orc_g2_to_uncompressed(&a, out);
    }
    return 0;
}

/* naive sum of base_i * exp_i (multiexp.rs:299-308 `naive_multiexp`), FullDensity */
int orc_naive_multiexp(const orc_bases_t* b, const uint64_t* exps, size_t n, uint8_t* out) {
    if (b->group == 1) {
        g1_t acc, p, t; g1_affine_t a; g1_identity(&acc);
        for (size_t i = 0; i < n; i++) {
            g1_from_affine(&p, (g1_affine_t*)b->pts + i); g1_mul(&t, &p, exps + 4 * i); g1_add(&acc, &acc, &t);
        }
        g1_to_affine(&a, &acc); orc_g1_to_uncompressed(&a, out);
    } else {
        g2_t acc, p, t; g2_affine_t a; g2_identity(&acc);
        for (size_t i = 0; i < n; i++) {
            g2_from_affine(&p, (g2_affine_t*)b->pts + i); g2_mul(&t, &p, exps + 4 * i); g2_add(&acc, &acc, &t);
        }
        g2_to_affine(&a, &acc); orc_g2_to_uncompressed(&a, out);
    }
    return 0;
}

/* generator * k as uncompressed bytes, and sum_i k_i * s_i mod q (known-dlog expectations) */
void orc_g1_generator_mul(const uint64_t k[4], uint8_t out[96]) {
    g1_t g, r; g1_affine_t a;
    memcpy(g.x.l, G1_GEN_X, 48); memcpy(g.y.l, G1_GEN_Y, 48); fp_one(&g.z);
    g1_mul(&r, &g, k); g1_to_affine(&a, &r); orc_g1_to_uncompressed(&a, out);
}
void orc_fr_dot(const uint64_t* k, const uint64_t* s, size_t n, uint64_t out[4]) {
    /* canonical in, canonical out */
    fr_t acc; fr_zero(&acc);
    for (size_t i = 0; i < n; i++) {
        fr_t a, b, p;
        memcpy(a.l, k + 4 * i, 32); memcpy(b.l, s + 4 * i, 32);
        fr_to_mont(&a, &a); fr_to_mont(&b, &b);
        fr_mul(&p, &a, &b); fr_add(&acc, &acc, &p);
    }
    fr_from_mont(&acc, &acc);
    memcpy(out, acc.l, 32);
}

/* -------------------------------------------------------------------------------- domain */
static uint32_t bitreverse(uint32_t n, uint32_t l) {
    uint32_t r = 0;
    for (uint32_t i = 0; i < l; i++) { r = (r << 1) | (n & 1); n >>= 1; }
    return r;
}
static void fr_pow_u64(fr_t* r, const fr_t* a, uint64_t e) { fr_pow(r, a, &e, 1); }

static void serial_fft(fr_t* a, const fr_t* omega, uint32_t log_n) { /* domain.rs:272-314 */
    uint32_t n = 1u << log_n;
    for (uint32_t k = 0; k < n; k++) {
        uint32_t rk = bitreverse(k, log_n);
        if (k < rk) { fr_t t = a[rk]; a[rk] = a[k]; a[k] = t; }
    }
    uint32_t m = 1;
    for (uint32_t s = 0; s < log_n; s++) {
        fr_t w_m; fr_pow_u64(&w_m, omega, n / (2 * m));
        for (uint32_t k = 0; k < n; k += 2 * m) {
            fr_t w; fr_one(&w);
            for (uint32_t j = 0; j < m; j++) {
                fr_t t; fr_mul(&t, &a[k + j + m], &w);
                fr_t tmp; fr_sub(&tmp, &a[k + j], &t);
                a[k + j + m] = tmp;
                fr_add(&a[k + j], &a[k + j], &t);
                fr_mul(&w, &w, &w_m);
            }
        }
        m *= 2;
    }
}

typedef struct { const fr_t* a; fr_t** tmp; const fr_t* omega; fr_t new_omega; uint32_t log_n, log_cpus; } pfft_t;
static void pfft_task(void* arg, size_t j) { /* domain.rs:331-356 */
    pfft_t* p = (pfft_t*)arg;
    uint32_t log_new_n = p->log_n - p->log_cpus;
    uint32_t num_cpus = 1u << p->log_cpus;
    fr_t omega_j, omega_step, elt;
    fr_pow_u64(&omega_j, p->omega, j);
    fr_pow_u64(&omega_step, p->omega, (uint64_t)j << log_new_n);
    fr_one(&elt);
    fr_t* tmp = p->tmp[j];
    size_t mask = ((size_t)1 << p->log_n) - 1;
    for (size_t i = 0; i < ((size_t)1 << log_new_n); i++) {
        fr_t acc; fr_zero(&acc);
        for (uint32_t s = 0; s < num_cpus; s++) {
            size_t idx = (i + ((size_t)s << log_new_n)) & mask;
            fr_t t; fr_mul(&t, &p->a[idx], &elt);
            fr_add(&acc, &acc, &t);
            fr_mul(&elt, &elt, &omega_step);
        }
        tmp[i] = acc;
        fr_mul(&elt, &elt, &omega_j);
    }
    serial_fft(tmp, &p->new_omega, log_new_n);
}
typedef struct { fr_t* a; fr_t** tmp; size_t chunk, n; uint32_t log_cpus; } pcopy_t;
static void pcopy_task(void* arg, size_t c) { /* domain.rs:358-371 */
    pcopy_t* p = (pcopy_t*)arg;
    size_t lo = c * p->chunk, hi = lo + p->chunk;
    if (hi > p->n) hi = p->n;
    size_t mask = ((size_t)1 << p->log_cpus) - 1;
    for (size_t idx = lo; idx < hi; idx++) p->a[idx] = p->tmp[idx & mask][idx >> p->log_cpus];
}
static void parallel_fft(fr_t* a, const fr_t* omega, uint32_t log_n, uint32_t log_cpus, int threads) {
    uint32_t num_cpus = 1u << log_cpus, log_new_n = log_n - log_cpus;
    fr_t** tmp = (fr_t**)malloc(sizeof(fr_t*) * num_cpus);
    for (uint32_t j = 0; j < num_cpus; j++) tmp[j] = (fr_t*)malloc(sizeof(fr_t) << log_new_n);
    pfft_t p = {a, tmp, omega, {{0}}, log_n, log_cpus};
    fr_pow_u64(&p.new_omega, omega, num_cpus);
    run_tasks(pfft_task, &p, num_cpus, threads);
    size_t n = (size_t)1 << log_n;
    size_t chunk = n < (size_t)threads ? 1 : n / threads; /* Worker::scope, multicore.rs:78-91 */
    pcopy_t pc = {a, tmp, chunk, n, log_cpus};
    run_tasks(pcopy_task, &pc, (n + chunk - 1) / chunk, threads);
    for (uint32_t j = 0; j < num_cpus; j++) free(tmp[j]);
    free(tmp);
}
static void best_fft(fr_t* a, const fr_t* omega, uint32_t log_n, int threads) { /* domain.rs:261-269 */
    uint32_t log_cpus = log2_floor((size_t)(threads < 1 ? 1 : threads));
    if (log_n <= log_cpus) serial_fft(a, omega, log_n);
    else parallel_fft(a, omega, log_n, log_cpus, threads);
}

typedef struct { fr_t* a; const fr_t* b; fr_t k; fr_t g; size_t chunk, n; int what; } pw_t;
static void pw_task(void* arg, size_t c) {
    pw_t* p = (pw_t*)arg;
    size_t lo = c * p->chunk, hi = lo + p->chunk;
    if (hi > p->n) hi = p->n;
    if (p->what == 0) for (size_t i = lo; i < hi; i++) fr_mul(&p->a[i], &p->a[i], &p->k);        /* scale */
    else if (p->what == 1) for (size_t i = lo; i < hi; i++) fr_mul(&p->a[i], &p->a[i], &p->b[i]); /* mul_assign */
    else if (p->what == 2) for (size_t i = lo; i < hi; i++) fr_sub(&p->a[i], &p->a[i], &p->b[i]); /* sub_assign */
    else if (p->what == 3) { /* distribute_powers, domain.rs:101-113 */
        fr_t u; fr_pow_u64(&u, &p->g, lo);
        for (size_t i = lo; i < hi; i++) { fr_mul(&p->a[i], &p->a[i], &u); fr_mul(&u, &u, &p->g); }
    } else for (size_t i = lo; i < hi; i++) fr_from_mont(&p->a[i], &p->a[i]);                    /* to_le_bits */
}
static void pointwise(int what, fr_t* a, const fr_t* b, const fr_t* k, const fr_t* g, size_t n, int threads) {
    if (!n) return;
    size_t chunk = n < (size_t)threads ? 1 : n / threads;
    pw_t p; memset(&p, 0, sizeof(p));
    p.a = a; p.b = b; p.chunk = chunk; p.n = n; p.what = what;
    if (k) p.k = *k;
    if (g) p.g = *g;
    run_tasks(pw_task, &p, (n + chunk - 1) / chunk, threads);
}

typedef struct { fr_t omega, omegainv, geninv, minv, gen; uint32_t exp; size_t m; } domain_t;
static void domain_init(domain_t* d, uint32_t exp) { /* from_coeffs, domain.rs:62-77 */
    d->exp = exp; d->m = (size_t)1 << exp;
    memcpy(d->omega.l, FR_ROOT, 32);
    for (uint32_t i = exp; i < 32; i++) fr_sqr(&d->omega, &d->omega);
    fr_inv(&d->omegainv, &d->omega);
    fr_from_u64(&d->gen, 7);
    fr_inv(&d->geninv, &d->gen);
    fr_t m; fr_from_u64(&m, (uint64_t)d->m);
    fr_inv(&d->minv, &m);
}
static void dom_fft(const domain_t* d, fr_t* a, int threads) { best_fft(a, &d->omega, d->exp, threads); }
static void dom_ifft(const domain_t* d, fr_t* a, int threads) { /* domain.rs:85-99 */
    best_fft(a, &d->omegainv, d->exp, threads);
    pointwise(0, a, NULL, &d->minv, NULL, d->m, threads);
}
static void dom_coset_fft(const domain_t* d, fr_t* a, int threads) { /* :115-118 */
    pointwise(3, a, NULL, NULL, &d->gen, d->m, threads);
    dom_fft(d, a, threads);
}
static void dom_icoset_fft(const domain_t* d, fr_t* a, int threads) { /* :120-125 */
    dom_ifft(d, a, threads);
    pointwise(3, a, NULL, NULL, &d->geninv, d->m, threads);
}
static void dom_divide_by_z_on_coset(const domain_t* d, fr_t* a, int threads) { /* :129-151 */
    fr_t z, one, i;
    fr_pow_u64(&z, &d->gen, (uint64_t)d->m);
    fr_one(&one); fr_sub(&z, &z, &one);
    fr_inv(&i, &z);
    pointwise(0, a, NULL, &i, NULL, d->m, threads);
}

/* In-place transform of m = 2^log_m Montgomery coefficients; op as BMPC_FFT.. (0..3) */
int orc_ntt(uint64_t* coeffs, uint32_t log_m, int op, int threads) {
    if (log_m >= 32) return 3;
    domain_t d; domain_init(&d, log_m);
    fr_t* a = (fr_t*)coeffs;
    switch (op) {
        case 0: dom_fft(&d, a, threads); break;
        case 1: dom_ifft(&d, a, threads); break;
        case 2: dom_coset_fft(&d, a, threads); break;
        case 3: dom_icoset_fft(&d, a, threads); break;
        default: return 6;
    }
    return 0;
}

/* prover.rs:210-231.  a, b, c: len Montgomery evaluations; out: (m-1) canonical scalars. */
int orc_h_coefficients(const uint64_t* a_in, const uint64_t* b_in, const uint64_t* c_in, size_t len,
                       uint64_t* out, size_t* out_len, int threads) {
    size_t m = 1; uint32_t exp = 0;
    while (m < len) { m *= 2; exp++; if (exp >= 32) return 3; }
    domain_t d; domain_init(&d, exp);
    fr_t* p[3];
    const uint64_t* src[3] = {a_in, b_in, c_in};
    for (int k = 0; k < 3; k++) {
        p[k] = (fr_t*)calloc(m, sizeof(fr_t));
        memcpy(p[k], src[k], len * 32);
        dom_ifft(&d, p[k], threads);
        dom_coset_fft(&d, p[k], threads);
    }
    pointwise(1, p[0], p[1], NULL, NULL, m, threads);
    pointwise(2, p[0], p[2], NULL, NULL, m, threads);
    dom_divide_by_z_on_coset(&d, p[0], threads);
    dom_icoset_fft(&d, p[0], threads);
    /* `to_le_bits` map -- serial in the reference (prover.rs:231 "TODO: parallelize") */
    for (size_t i = 0; i + 1 < m; i++) fr_from_mont(&p[0][i], &p[0][i]);
    memcpy(out, p[0], (m - 1) * 32);
    *out_len = m - 1;
    for (int k = 0; k < 3; k++) free(p[k]);
    return 0;
}

/* ------------------------------------------------------------------------- create_proof
 * prover.rs:206-350.  The eight multiexps are in flight together in the reference (rayon
 * tasks, one per window each); here all their window regions form one task list. */
typedef struct {
    const orc_bases_t *h, *l, *a, *b_g1, *b_g2;
    uint8_t alpha_g1[96], beta_g1[96], beta_g2[192], delta_g1[96], delta_g2[192];
} orc_params_t;

typedef struct { mexp_job_t job; size_t nwin; } mexp_slot_t;
typedef struct { mexp_slot_t* slots; int nslots; size_t* first; } multi_t;
static void multi_task(void* arg, size_t idx) {
    multi_t* m = (multi_t*)arg;
    int s = 0;
    while (s + 1 < m->nslots && idx >= m->first[s + 1]) s++;
    mexp_task(&m->slots[s].job, idx - m->first[s]);
}
static void compress_g1(const g1_affine_t* a, uint8_t out[48]) {
    if (a->inf) { memset(out, 0, 48); out[0] = 0xc0; return; }
    fp_to_be(&a->x, out); out[0] |= 0x80;
    fp_t y, ny; fp_from_mont(&y, &a->y); fp_neg(&ny, &a->y); fp_from_mont(&ny, &ny);
    int larger = 0;
    for (int i = 5; i >= 0; i--) { if (y.l[i] != ny.l[i]) { larger = y.l[i] > ny.l[i]; break; } }
    if (larger) out[0] |= 0x20;
}
static int fp_lex_largest(const fp_t* m) {
    fp_t y, ny; fp_from_mont(&y, m); fp_neg(&ny, m); fp_from_mont(&ny, &ny);
    for (int i = 5; i >= 0; i--) if (y.l[i] != ny.l[i]) return y.l[i] > ny.l[i];
    return 0;
}
static void compress_g2(const g2_affine_t* a, uint8_t out[96]) {
    if (a->inf) { memset(out, 0, 96); out[0] = 0xc0; return; }
    fp_to_be(&a->x.c1, out); fp_to_be(&a->x.c0, out + 48); out[0] |= 0x80;
    int larger = fp_is_zero(&a->y.c1) ? fp_lex_largest(&a->y.c0) : fp_lex_largest(&a->y.c1);
    if (larger) out[0] |= 0x20;
}

int orc_create_proof(const orc_params_t* P, const uint64_t* a, const uint64_t* b, const uint64_t* c,
                     size_t num_constraints, const uint64_t* inputs_mont, size_t ni,
                     const uint64_t* aux_mont, size_t na, const uint64_t* a_aux_density,
                     const uint64_t* b_input_density, const uint64_t* b_aux_density,
                     const uint64_t r_mont[4], const uint64_t s_mont[4], int threads, uint8_t proof[192]) {
    size_t m = 1;
    while (m < num_constraints) m *= 2;
    uint64_t* h_exps = (uint64_t*)malloc((m ? m : 1) * 32);
    size_t h_len;
    int st = orc_h_coefficients(a, b, c, num_constraints, h_exps, &h_len, threads);
    if (st) { free(h_exps); return st; }
    fr_t* in = (fr_t*)malloc((ni ? ni : 1) * 32);
    fr_t* ax = (fr_t*)malloc((na ? na : 1) * 32);
    for (size_t i = 0; i < ni; i++) { memcpy(&in[i], inputs_mont + 4 * i, 32); fr_from_mont(&in[i], &in[i]); } /* :237-250 */
    for (size_t i = 0; i < na; i++) { memcpy(&ax[i], aux_mont + 4 * i, 32); fr_from_mont(&ax[i], &ax[i]); }
    size_t b_in_total = 0;
    for (size_t i = 0; i < ni; i++) b_in_total += (b_input_density[i >> 6] >> (i & 63)) & 1;

    g1_t r_g1[6]; g2_t r_g2[2];
    struct { const orc_bases_t* b; size_t start; const uint64_t* e; size_t n; const uint64_t* d; } q[8] = {
        {P->a, 0, (uint64_t*)in, ni, NULL},                  /* a_inputs */
        {P->a, ni, (uint64_t*)ax, na, a_aux_density},        /* a_aux */
        {P->b_g1, 0, (uint64_t*)in, ni, b_input_density},    /* b_g1_inputs */
        {P->b_g1, b_in_total, (uint64_t*)ax, na, b_aux_density},
        {P->b_g2, 0, (uint64_t*)in, ni, b_input_density},
        {P->b_g2, b_in_total, (uint64_t*)ax, na, b_aux_density},
        {P->h, 0, h_exps, h_len, NULL},
        {P->l, 0, (uint64_t*)ax, na, NULL},
    };
    mexp_slot_t slots[8];
    size_t first[9];
    first[0] = 0;
    for (int k = 0; k < 8; k++) {
        uint32_t cw = orc_window_size(q[k].n);
        size_t nwin = (255 + cw - 1) / cw;
        int grp = q[k].b->group;
        slots[k].nwin = nwin;
        slots[k].job.group = grp; slots[k].job.bases = q[k].b->pts; slots[k].job.nbases = q[k].b->n;
        slots[k].job.start = q[k].start; slots[k].job.exps = q[k].e; slots[k].job.n = q[k].n;
        slots[k].job.density = q[k].d; slots[k].job.c = cw;
        slots[k].job.parts = malloc(nwin * (grp == 1 ? sizeof(g1_t) : sizeof(g2_t)));
        slots[k].job.errs = (int*)calloc(nwin, sizeof(int));
        first[k + 1] = first[k] + nwin;
    }
    multi_t mt = {slots, 8, first};
    run_tasks(multi_task, &mt, first[8], threads);
    int statuses[8];
    g1_t* g1out[8] = {&r_g1[0], &r_g1[1], &r_g1[2], &r_g1[3], NULL, NULL, &r_g1[4], &r_g1[5]};
    for (int k = 0; k < 8; k++) { /* the fold of multiexp.rs:244-249 */
        uint32_t cw = slots[k].job.c;
        statuses[k] = 0;
        if (slots[k].job.group == 1) {
            g1_t acc; g1_identity(&acc);
            for (size_t w = slots[k].nwin; w-- > 0;) {
                if (slots[k].job.errs[w]) { statuses[k] = slots[k].job.errs[w]; break; }
                for (uint32_t t = 0; t < cw; t++) g1_dbl(&acc, &acc);
                g1_add(&acc, &acc, (g1_t*)slots[k].job.parts + w);
            }
            *g1out[k] = acc;
        } else {
            g2_t acc; g2_identity(&acc);
            for (size_t w = slots[k].nwin; w-- > 0;) {
                if (slots[k].job.errs[w]) { statuses[k] = slots[k].job.errs[w]; break; }
                for (uint32_t t = 0; t < cw; t++) g2_dbl(&acc, &acc);
                g2_add(&acc, &acc, (g2_t*)slots[k].job.parts + w);
            }
            r_g2[k - 4] = acc;
        }
        free(slots[k].job.parts); free(slots[k].job.errs);
    }
    free(h_exps); free(in); free(ax);
    if ((P->delta_g1[0] & 0x40) || (P->delta_g2[0] & 0x40)) return 1; /* prover.rs:309-313 */
    for (int k = 0; k < 8; k++) if (statuses[k]) return statuses[k];

    g1_affine_t alpha_a, beta1_a, delta1_a; g2_affine_t beta2_a, delta2_a;
    g1_from_uncompressed(&alpha_a, P->alpha_g1); g1_from_uncompressed(&beta1_a, P->beta_g1);
    g1_from_uncompressed(&delta1_a, P->delta_g1);
    g2_from_uncompressed(&beta2_a, P->beta_g2); g2_from_uncompressed(&delta2_a, P->delta_g2);
    g1_t alpha, beta1, delta1; g2_t beta2, delta2;
    g1_from_affine(&alpha, &alpha_a); g1_from_affine(&beta1, &beta1_a); g1_from_affine(&delta1, &delta1_a);
    g2_from_affine(&beta2, &beta2_a); g2_from_affine(&delta2, &delta2_a);
    fr_t r, s, rs, rc, sc, rsc;
    memcpy(r.l, r_mont, 32); memcpy(s.l, s_mont, 32);
    fr_mul(&rs, &r, &s);
    fr_from_mont(&rc, &r); fr_from_mont(&sc, &s); fr_from_mont(&rsc, &rs);
    g1_t g_a, g_c, t; g2_t g_b;
    g1_mul(&g_a, &delta1, rc.l); g1_add(&g_a, &g_a, &alpha);            /* :315-316 */
    g2_mul(&g_b, &delta2, sc.l); g2_add(&g_b, &g_b, &beta2);            /* :317-318 */
    g1_mul(&g_c, &delta1, rsc.l);                                       /* :319-327 */
    g1_mul(&t, &alpha, sc.l); g1_add(&g_c, &g_c, &t);
    g1_mul(&t, &beta1, rc.l); g1_add(&g_c, &g_c, &t);
    g1_t a_answer; g1_add(&a_answer, &r_g1[0], &r_g1[1]);               /* :328-332 */
    g1_add(&g_a, &g_a, &a_answer);
    g1_mul(&t, &a_answer, sc.l); g1_add(&g_c, &g_c, &t);
    g1_t b1_answer; g1_add(&b1_answer, &r_g1[2], &r_g1[3]);             /* :334-337 */
    g2_t b2_answer; g2_add(&b2_answer, &r_g2[0], &r_g2[1]);
    g2_add(&g_b, &g_b, &b2_answer);                                     /* :339-343 */
    g1_mul(&t, &b1_answer, rc.l); g1_add(&g_c, &g_c, &t);
    g1_add(&g_c, &g_c, &r_g1[4]);
    g1_add(&g_c, &g_c, &r_g1[5]);
    g1_affine_t A, Cc; g2_affine_t B;
    g1_to_affine(&A, &g_a); g2_to_affine(&B, &g_b); g1_to_affine(&Cc, &g_c);
    compress_g1(&A, proof); compress_g2(&B, proof + 48); compress_g1(&Cc, proof + 144);
    return 0;
}

int orc_hardware_threads(void) {
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) return CPU_COUNT(&set);
    return 1;
}
