/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the BLS12-381 field arithmetic.
 *
 * The reference's arithmetic is the un-vendored crate bls12_381 0.6.0 (Cargo.lock:96-99):
 * 64-bit-limb Montgomery form, R = 2^256 (Scalar) / 2^384 (Fp), fully reduced.  This file
 * restates that published representation (same limbs bit-for-bit) so that outputs can be
 * compared byte-for-byte; the Fp modulus and INV also appear in the reference itself at
 * src/gt_bytes.rs:20-30.  Not product code: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference leg may link or load it.
 */
#ifndef ORC_FIELD_H
#define ORC_FIELD_H
#include <stdint.h>
#include <string.h>

typedef unsigned __int128 u128;

typedef struct { uint64_t l[4]; } fr_t;
typedef struct { uint64_t l[6]; } fp_t;
typedef struct { fp_t c0, c1; } fp2_t;

static const uint64_t FR_MOD[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL,
                                   0x73eda753299d7d48ULL};
static const uint64_t FR_INV = 0xfffffffeffffffffULL;
static const uint64_t FR_R[4] = {0x00000001fffffffeULL, 0x5884b7fa00034802ULL, 0x998c4fefecbc4ff5ULL,
                                 0x1824b159acc5056fULL};
static const uint64_t FR_R2[4] = {0xc999e990f3f29c6dULL, 0x2b6cedcb87925c23ULL, 0x05d314967254398fULL,
                                  0x0748d9d99f59ff11ULL};
/* root_of_unity = 7^((q-1)/2^32), Montgomery limbs */
static const uint64_t FR_ROOT[4] = {0xb9b58d8c5f0e466aULL, 0x5b1b4c801819d7ecULL, 0x0af53ae352a31e64ULL,
                                    0x5bf3adda19e9b27bULL};

static const uint64_t FP_MOD[6] = {0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                                   0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL};
static const uint64_t FP_INV = 0x89f3fffcfffcfffdULL;
static const uint64_t FP_R[6] = {0x760900000002fffdULL, 0xebf4000bc40c0002ULL, 0x5f48985753c758baULL,
                                 0x77ce585370525745ULL, 0x5c071a97a256ec6dULL, 0x15f65ec3fa80e493ULL};
static const uint64_t FP_R2[6] = {0xf4df1f341c341746ULL, 0x0a76e6a609d104f1ULL, 0x8de5476c4c95b6d5ULL,
                                  0x67eb88a9939d83c0ULL, 0x9a793e85b519952dULL, 0x11988fe592cae3aaULL};

#define ORC_DEFINE_FIELD(T, N, MOD, INV, ONE, PFX)                                              \
    static inline int PFX##_is_zero(const T* a) {                                               \
        uint64_t acc = 0;                                                                       \
        for (int i = 0; i < N; i++) acc |= a->l[i];                                             \
        return acc == 0;                                                                        \
    }                                                                                           \
    static inline int PFX##_eq(const T* a, const T* b) {                                        \
        uint64_t acc = 0;                                                                       \
        for (int i = 0; i < N; i++) acc |= a->l[i] ^ b->l[i];                                   \
        return acc == 0;                                                                        \
    }                                                                                           \
    static inline void PFX##_zero(T* r) { memset(r, 0, sizeof(T)); }                            \
    static inline void PFX##_one(T* r) { memcpy(r->l, ONE, sizeof(T)); }                        \
    /* r = a - MOD if a >= MOD */                                                               \
    static inline void PFX##_reduce_once(T* r, const uint64_t* a, uint64_t carry) {             \
        uint64_t t[N];                                                                          \
        u128 b = 0;                                                                             \
        for (int i = 0; i < N; i++) {                                                           \
            u128 d = (u128)a[i] - MOD[i] - (uint64_t)b;                                         \
            t[i] = (uint64_t)d;                                                                 \
            b = (d >> 64) & 1;                                                                  \
        }                                                                                       \
        int ge = carry || !b;                                                                   \
        for (int i = 0; i < N; i++) r->l[i] = ge ? t[i] : a[i];                                 \
    }                                                                                           \
    static inline void PFX##_add(T* r, const T* a, const T* b) {                                \
        uint64_t t[N];                                                                          \
        u128 c = 0;                                                                             \
        for (int i = 0; i < N; i++) {                                                           \
            c += (u128)a->l[i] + b->l[i];                                                       \
            t[i] = (uint64_t)c;                                                                 \
            c >>= 64;                                                                           \
        }                                                                                       \
        PFX##_reduce_once(r, t, (uint64_t)c);                                                   \
    }                                                                                           \
    static inline void PFX##_sub(T* r, const T* a, const T* b) {                                \
        uint64_t t[N];                                                                          \
        u128 br = 0;                                                                            \
        for (int i = 0; i < N; i++) {                                                           \
            u128 d = (u128)a->l[i] - b->l[i] - (uint64_t)br;                                    \
            t[i] = (uint64_t)d;                                                                 \
            br = (d >> 64) & 1;                                                                 \
        }                                                                                       \
        if (br) {                                                                               \
            u128 c = 0;                                                                         \
            for (int i = 0; i < N; i++) {                                                       \
                c += (u128)t[i] + MOD[i];                                                       \
                t[i] = (uint64_t)c;                                                             \
                c >>= 64;                                                                       \
            }                                                                                   \
        }                                                                                       \
        memcpy(r->l, t, sizeof(t));                                                             \
    }                                                                                           \
    static inline void PFX##_neg(T* r, const T* a) {                                            \
        T z;                                                                                    \
        PFX##_zero(&z);                                                                         \
        PFX##_sub(r, &z, a);                                                                    \
    }                                                                                           \
    static inline void PFX##_dbl(T* r, const T* a) { PFX##_add(r, a, a); }                      \
    /* Montgomery product (CIOS), fully reduced */                                              \
    static inline void PFX##_mul(T* r, const T* a, const T* b) {                                \
        uint64_t t[N + 2];                                                                      \
        memset(t, 0, sizeof(t));                                                                \
        for (int i = 0; i < N; i++) {                                                           \
            u128 c = 0;                                                                         \
            for (int j = 0; j < N; j++) {                                                       \
                c += (u128)a->l[j] * b->l[i] + t[j];                                            \
                t[j] = (uint64_t)c;                                                             \
                c >>= 64;                                                                       \
            }                                                                                   \
            c += t[N];                                                                          \
            t[N] = (uint64_t)c;                                                                 \
            t[N + 1] = (uint64_t)(c >> 64);                                                     \
            uint64_t m = t[0] * INV;                                                            \
            c = (u128)m * MOD[0] + t[0];                                                        \
            c >>= 64;                                                                           \
            for (int j = 1; j < N; j++) {                                                       \
                c += (u128)m * MOD[j] + t[j];                                                   \
                t[j - 1] = (uint64_t)c;                                                         \
                c >>= 64;                                                                       \
            }                                                                                   \
            c += t[N];                                                                          \
            t[N - 1] = (uint64_t)c;                                                             \
            t[N] = t[N + 1] + (uint64_t)(c >> 64);                                              \
        }                                                                                       \
        PFX##_reduce_once(r, t, t[N]);                                                          \
    }                                                                                           \
    static inline void PFX##_sqr(T* r, const T* a) { PFX##_mul(r, a, a); }                      \
    static inline void PFX##_pow(T* r, const T* a, const uint64_t* e, int words) {              \
        T acc;                                                                                  \
        PFX##_one(&acc);                                                                        \
        for (int i = words - 1; i >= 0; i--)                                                    \
            for (int b = 63; b >= 0; b--) {                                                     \
                PFX##_sqr(&acc, &acc);                                                          \
                if ((e[i] >> b) & 1) PFX##_mul(&acc, &acc, a);                                  \
            }                                                                                   \
        *r = acc;                                                                               \
    }                                                                                           \
    static inline void PFX##_inv(T* r, const T* a) {                                            \
        uint64_t e[N];                                                                          \
        memcpy(e, MOD, sizeof(e));                                                              \
        e[0] -= 2; /* both moduli end in ...01 / ...ab: no borrow */                            \
        PFX##_pow(r, a, e, N);                                                                  \
    }                                                                                           \
    static inline void PFX##_from_mont(T* r, const T* a) {                                      \
        T o;                                                                                    \
        PFX##_zero(&o);                                                                         \
        o.l[0] = 1;                                                                             \
        PFX##_mul(r, a, &o);                                                                    \
    }

ORC_DEFINE_FIELD(fr_t, 4, FR_MOD, FR_INV, FR_R, fr)
ORC_DEFINE_FIELD(fp_t, 6, FP_MOD, FP_INV, FP_R, fp)

static inline void fr_to_mont(fr_t* r, const fr_t* a) {
    fr_t r2;
    memcpy(r2.l, FR_R2, 32);
    fr_mul(r, a, &r2);
}
static inline void fp_to_mont(fp_t* r, const fp_t* a) {
    fp_t r2;
    memcpy(r2.l, FP_R2, 48);
    fp_mul(r, a, &r2);
}
static inline void fr_from_u64(fr_t* r, uint64_t v) {
    fr_t t;
    fr_zero(&t);
    t.l[0] = v;
    fr_to_mont(r, &t);
}

/* ---- Fp2 = Fp[u]/(u^2+1) ------------------------------------------------------------ */
static inline int fp2_is_zero(const fp2_t* a) { return fp_is_zero(&a->c0) && fp_is_zero(&a->c1); }
static inline int fp2_eq(const fp2_t* a, const fp2_t* b) { return fp_eq(&a->c0, &b->c0) && fp_eq(&a->c1, &b->c1); }
static inline void fp2_zero(fp2_t* r) { memset(r, 0, sizeof(*r)); }
static inline void fp2_one(fp2_t* r) { fp_one(&r->c0); fp_zero(&r->c1); }
static inline void fp2_add(fp2_t* r, const fp2_t* a, const fp2_t* b) { fp_add(&r->c0, &a->c0, &b->c0); fp_add(&r->c1, &a->c1, &b->c1); }
static inline void fp2_sub(fp2_t* r, const fp2_t* a, const fp2_t* b) { fp_sub(&r->c0, &a->c0, &b->c0); fp_sub(&r->c1, &a->c1, &b->c1); }
static inline void fp2_neg(fp2_t* r, const fp2_t* a) { fp_neg(&r->c0, &a->c0); fp_neg(&r->c1, &a->c1); }
static inline void fp2_dbl(fp2_t* r, const fp2_t* a) { fp2_add(r, a, a); }
static inline void fp2_mul(fp2_t* r, const fp2_t* a, const fp2_t* b) {
    fp_t t0, t1, s0, s1, t2;
    fp_mul(&t0, &a->c0, &b->c0);
    fp_mul(&t1, &a->c1, &b->c1);
    fp_add(&s0, &a->c0, &a->c1);
    fp_add(&s1, &b->c0, &b->c1);
    fp_mul(&t2, &s0, &s1);
    fp_sub(&r->c0, &t0, &t1);
    fp_sub(&t2, &t2, &t0);
    fp_sub(&r->c1, &t2, &t1);
}
static inline void fp2_sqr(fp2_t* r, const fp2_t* a) { fp2_mul(r, a, a); }
static inline void fp2_inv(fp2_t* r, const fp2_t* a) {
    fp_t n, t;
    fp_sqr(&n, &a->c0);
    fp_sqr(&t, &a->c1);
    fp_add(&n, &n, &t);
    fp_inv(&n, &n);
    fp_mul(&r->c0, &a->c0, &n);
    fp_mul(&t, &a->c1, &n);
    fp_neg(&r->c1, &t);
}

#endif
