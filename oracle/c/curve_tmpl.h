/* TEST INFRASTRUCTURE ONLY -- group law "template", included once per group with
 *   FE      coordinate field type            FE_(op)   field function prefix macro
 *   PT      projective point type name       AF        affine point type name
 *   GP_(f)  group function prefix macro      MUL_3B    r = 3b * a for this curve
 *
 * Restates the published complete formulas the third-party crate bls12_381 0.6.0 uses for
 * G1Projective / G2Projective (Renes-Costello-Batina 2015, "Complete addition formulas for
 * prime order elliptic curves", Algorithms 7-9, a = 0): homogeneous projective coordinates,
 * identity (0 : 1 : 0).  The reference reaches them through `add_assign` / `double` at
 * src/multiexp.rs:39,217,231-232,248.  Using the same formulas keeps the timed CPU baseline's
 * cost per group operation comparable with the Rust reference.
 */

typedef struct { FE x, y; int inf; } AF;
typedef struct { FE x, y, z; } PT;

static inline void GP_(identity)(PT* r) { FE_(zero)(&r->x); FE_(one)(&r->y); FE_(zero)(&r->z); }
static inline int GP_(is_identity)(const PT* p) { return FE_(is_zero)(&p->z); }

/* Algorithm 7: complete addition */
static inline void GP_(add)(PT* r, const PT* p, const PT* q) {
    FE t0, t1, t2, t3, t4, x3, y3, z3;
    FE_(mul)(&t0, &p->x, &q->x);
    FE_(mul)(&t1, &p->y, &q->y);
    FE_(mul)(&t2, &p->z, &q->z);
    FE_(add)(&t3, &p->x, &p->y);
    FE_(add)(&t4, &q->x, &q->y);
    FE_(mul)(&t3, &t3, &t4);
    FE_(add)(&t4, &t0, &t1);
    FE_(sub)(&t3, &t3, &t4);
    FE_(add)(&t4, &p->y, &p->z);
    FE_(add)(&x3, &q->y, &q->z);
    FE_(mul)(&t4, &t4, &x3);
    FE_(add)(&x3, &t1, &t2);
    FE_(sub)(&t4, &t4, &x3);
    FE_(add)(&x3, &p->x, &p->z);
    FE_(add)(&y3, &q->x, &q->z);
    FE_(mul)(&x3, &x3, &y3);
    FE_(add)(&y3, &t0, &t2);
    FE_(sub)(&y3, &x3, &y3);
    FE_(add)(&x3, &t0, &t0);
    FE_(add)(&t0, &x3, &t0);
    MUL_3B(&t2, &t2);
    FE_(add)(&z3, &t1, &t2);
    FE_(sub)(&t1, &t1, &t2);
    MUL_3B(&y3, &y3);
    FE_(mul)(&x3, &t4, &y3);
    FE_(mul)(&t2, &t3, &t1);
    FE_(sub)(&x3, &t2, &x3);
    FE_(mul)(&y3, &y3, &t0);
    FE_(mul)(&t1, &t1, &z3);
    FE_(add)(&y3, &t1, &y3);
    FE_(mul)(&t0, &t0, &t3);
    FE_(mul)(&z3, &z3, &t4);
    FE_(add)(&z3, &z3, &t0);
    r->x = x3; r->y = y3; r->z = z3;
}

/* Algorithm 8: mixed addition (q affine, not the identity) */
static inline void GP_(add_mixed)(PT* r, const PT* p, const AF* q) {
    if (q->inf) { *r = *p; return; }
    FE t0, t1, t2, t3, t4, x3, y3, z3;
    FE_(mul)(&t0, &p->x, &q->x);
    FE_(mul)(&t1, &p->y, &q->y);
    FE_(add)(&t3, &q->x, &q->y);
    FE_(add)(&t4, &p->x, &p->y);
    FE_(mul)(&t3, &t3, &t4);
    FE_(add)(&t4, &t0, &t1);
    FE_(sub)(&t3, &t3, &t4);
    FE_(mul)(&t4, &q->y, &p->z);
    FE_(add)(&t4, &t4, &p->y);
    FE_(mul)(&y3, &q->x, &p->z);
    FE_(add)(&y3, &y3, &p->x);
    FE_(add)(&x3, &t0, &t0);
    FE_(add)(&t0, &x3, &t0);
    MUL_3B(&t2, &p->z);
    FE_(add)(&z3, &t1, &t2);
    FE_(sub)(&t1, &t1, &t2);
    MUL_3B(&y3, &y3);
    FE_(mul)(&x3, &t4, &y3);
    FE_(mul)(&t2, &t3, &t1);
    FE_(sub)(&x3, &t2, &x3);
    FE_(mul)(&y3, &y3, &t0);
    FE_(mul)(&t1, &t1, &z3);
    FE_(add)(&y3, &t1, &y3);
    FE_(mul)(&t0, &t0, &t3);
    FE_(mul)(&z3, &z3, &t4);
    FE_(add)(&z3, &z3, &t0);
    r->x = x3; r->y = y3; r->z = z3;
}

/* Algorithm 9: doubling */
static inline void GP_(dbl)(PT* r, const PT* p) {
    FE t0, t1, t2, x3, y3, z3;
    FE_(sqr)(&t0, &p->y);
    FE_(add)(&z3, &t0, &t0);
    FE_(add)(&z3, &z3, &z3);
    FE_(add)(&z3, &z3, &z3);
    FE_(mul)(&t1, &p->y, &p->z);
    FE_(sqr)(&t2, &p->z);
    MUL_3B(&t2, &t2);
    FE_(mul)(&x3, &t2, &z3);
    FE_(add)(&y3, &t0, &t2);
    FE_(mul)(&z3, &t1, &z3);
    FE_(add)(&t1, &t2, &t2);
    FE_(add)(&t2, &t1, &t2);
    FE_(sub)(&t0, &t0, &t2);
    FE_(mul)(&y3, &t0, &y3);
    FE_(add)(&y3, &x3, &y3);
    FE_(mul)(&t1, &p->x, &p->y);
    FE_(mul)(&x3, &t0, &t1);
    FE_(add)(&x3, &x3, &x3);
    r->x = x3; r->y = y3; r->z = z3;
}

static inline void GP_(to_affine)(AF* r, const PT* p) {
    if (GP_(is_identity)(p)) { FE_(zero)(&r->x); FE_(zero)(&r->y); r->inf = 1; return; }
    FE zi;
    FE_(inv)(&zi, &p->z);
    FE_(mul)(&r->x, &p->x, &zi);
    FE_(mul)(&r->y, &p->y, &zi);
    r->inf = 0;
}

static inline void GP_(from_affine)(PT* r, const AF* a) {
    if (a->inf) { GP_(identity)(r); return; }
    r->x = a->x; r->y = a->y; FE_(one)(&r->z);
}

/* p * k, k = 4 x u64 canonical little-endian (double-and-add; `Mul<Scalar>` in bls12_381) */
static inline void GP_(mul)(PT* r, const PT* p, const uint64_t k[4]) {
    PT acc;
    GP_(identity)(&acc);
    for (int i = 3; i >= 0; i--)
        for (int b = 63; b >= 0; b--) {
            GP_(dbl)(&acc, &acc);
            if ((k[i] >> b) & 1) GP_(add)(&acc, &acc, p);
        }
    *r = acc;
}

/* ------------------------------------------------------------------------------------
 * multiexp restatement: src/multiexp.rs:159-281.  One call = one window region (`this`
 * closure, :173-236); the caller folds the regions (:244-249).
 * bases: array of AF (inf flag = identity), nbases entries; start = Source cursor.
 * Returns 0 ok, 1 UnexpectedIdentity, 2 UnexpectedEof (first error in scan order). */
static int GP_(multiexp_region)(const AF* bases, size_t nbases, size_t start, const uint64_t* exps,
                                size_t n, const uint64_t* density, uint32_t skip, uint32_t c, PT* out) {
    PT acc;
    GP_(identity)(&acc);
    size_t nbuckets = ((size_t)1 << c) - 1;
    PT* buckets = (PT*)malloc(nbuckets * sizeof(PT));
    for (size_t i = 0; i < nbuckets; i++) GP_(identity)(&buckets[i]);
    size_t cur = start;
    int handle_trivial = skip == 0;
    int err = 0;
    for (size_t i = 0; i < n && !err; i++) {
        if (density && !((density[i >> 6] >> (i & 63)) & 1)) continue; /* :192 */
        const uint64_t* e = exps + 4 * i;
        int is_zero = (e[0] | e[1] | e[2] | e[3]) == 0;
        int is_one = e[0] == 1 && (e[1] | e[2] | e[3]) == 0;
        uint64_t d = 0;
        int consume = 0; /* 0 = skip(1), 1 = next() into acc, 2 = next() into bucket */
        if (is_zero) consume = 0;                   /* :199-200 */
        else if (is_one) consume = handle_trivial;  /* :201-206 */
        else {
            /* bits [skip, skip+c) of the 256-bit view; bits >= 256 are zero (:208-214) */
            uint32_t w = skip >> 6, sh = skip & 63;
            uint64_t lo = w < 4 ? e[w] : 0, hi = (w + 1) < 4 ? e[w + 1] : 0;
            d = sh ? ((lo >> sh) | (hi << (64 - sh))) : lo;
            d &= ((uint64_t)1 << c) - 1;
            consume = d ? 2 : 0;
        }
        if (cur >= nbases) { err = 2; break; }       /* next()/skip(): :55-61,74-80 */
        if (consume) {
            if (bases[cur].inf) { err = 1; break; }  /* next(): :63-65 */
            if (consume == 1) GP_(add_mixed)(&acc, &acc, &bases[cur]);
            else GP_(add_mixed)(&buckets[d - 1], &buckets[d - 1], &bases[cur]); /* :217 */
        }
        cur++;
    }
    if (!err) { /* summation by parts, :229-233 */
        PT running;
        GP_(identity)(&running);
        for (size_t i = nbuckets; i-- > 0;) {
            GP_(add)(&running, &running, &buckets[i]);
            GP_(add)(&acc, &acc, &running);
        }
        *out = acc;
    }
    free(buckets);
    return err;
}
