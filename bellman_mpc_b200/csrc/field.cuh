// Multi-limb Montgomery field arithmetic for BLS12-381 Fr (8 x u32) and Fp (12 x u32).
//
// Replaces the third-party `bls12_381::Scalar` / `Fp` arithmetic that the reference
// reaches through ff::Field at src/domain.rs:250-257,294-306 (mul/add/sub_assign in the
// butterfly) and through group ops at src/multiexp.rs:217,231-232,248.
//
// Representation is bit-compatible with bls12_381: little-endian limbs, Montgomery form
// with R = 2^256 (Fr) / 2^384 (Fp), every value fully reduced to [0, p).  A u32[8] here
// is the same 32 bytes as the reference's [u64; 4] on a little-endian host.
//
// Arithmetic is written as PTX carry chains (mad.lo.cc / madc.hi.cc ...), with the
// products of even- and odd-indexed limbs accumulated in two separate chains so every
// chain is a straight run of IMADs.  The same source compiles for the host with the
// carry flag emulated in a thread-local (used by tests/host_check to validate the
// algorithms bit-for-bit against the big-integer oracle without a GPU).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BMPC_HD __host__ __device__ __forceinline__
#define BMPC_D __device__ __forceinline__
// cold-path functions (inversion, exponentiation, full point addition ...) are kept out of
// line: inlining every 600-instruction Montgomery product into them makes device
// compilation explode and buys nothing where they run.
#define BMPC_COLD __host__ __device__ __noinline__
#else
#define BMPC_HD inline
#define BMPC_D inline
#define BMPC_COLD inline
#endif

namespace bmpc {

// ----------------------------------------------------------------------------- carry ops
#if defined(__CUDA_ARCH__)
BMPC_D uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
BMPC_D uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
BMPC_D uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
BMPC_D uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
BMPC_D uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
BMPC_D uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
BMPC_D uint32_t mul_lo(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
BMPC_D uint32_t mul_hi(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
BMPC_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
BMPC_D uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
BMPC_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
BMPC_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
BMPC_D uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
#else
// Host emulation of the PTX condition-code register (tests only).
inline uint32_t& cc_flag() { static thread_local uint32_t f = 0; return f; }
inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; cc_flag() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + cc_flag(); cc_flag() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + cc_flag(); }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; cc_flag() = (uint32_t)(t >> 63); return (uint32_t)t; }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - cc_flag(); cc_flag() = (uint32_t)(t >> 63); return (uint32_t)t; }
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - cc_flag(); }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(uint32_t)(a * b) + c; cc_flag() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (((uint64_t)a * b) >> 32) + c; cc_flag() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(uint32_t)(a * b) + c + cc_flag(); cc_flag() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (((uint64_t)a * b) >> 32) + c + cc_flag(); cc_flag() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return (uint32_t)((((uint64_t)a * b) >> 32) + c + cc_flag()); }
#endif

// -------------------------------------------------------------------- field parameters
// Constants re-derived by oracle/fields.py::self_check(); Fp modulus and INV also appear
// in the reference itself at src/gt_bytes.rs:20-30.
struct FrParams {
    static constexpr int N = 8;
    static constexpr uint32_t INV = 0xffffffffu;  // -q^-1 mod 2^32
    BMPC_HD static constexpr uint32_t mod(int i) {
        constexpr uint32_t m[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u,
                                   0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
        return m[i];
    }
    BMPC_HD static constexpr uint32_t one(int i) {  // R mod q
        constexpr uint32_t m[8] = {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau,
                                   0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u};
        return m[i];
    }
    BMPC_HD static constexpr uint32_t r2(int i) {  // R^2 mod q
        constexpr uint32_t m[8] = {0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu,
                                   0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u};
        return m[i];
    }
};

struct FpParams {
    static constexpr int N = 12;
    static constexpr uint32_t INV = 0xfffcfffdu;  // -p^-1 mod 2^32
    BMPC_HD static constexpr uint32_t mod(int i) {
        constexpr uint32_t m[12] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu,
                                    0xf6b0f624u, 0x6730d2a0u, 0xf38512bfu, 0x64774b84u,
                                    0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
        return m[i];
    }
    BMPC_HD static constexpr uint32_t one(int i) {  // R mod p
        constexpr uint32_t m[12] = {0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu,
                                    0x53c758bau, 0x5f489857u, 0x70525745u, 0x77ce5853u,
                                    0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u};
        return m[i];
    }
    BMPC_HD static constexpr uint32_t r2(int i) {  // R^2 mod p
        constexpr uint32_t m[12] = {0x1c341746u, 0xf4df1f34u, 0x09d104f1u, 0x0a76e6a6u,
                                    0x4c95b6d5u, 0x8de5476cu, 0x939d83c0u, 0x67eb88a9u,
                                    0xb519952du, 0x9a793e85u, 0x92cae3aau, 0x11988fe5u};
        return m[i];
    }
};

// ------------------------------------------------------------------------------ Field
template <class P>
struct Field {
    static constexpr int N = P::N;
    uint32_t l[N];

    BMPC_HD static Field zero() {
        Field r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = 0;
        return r;
    }
    BMPC_HD static Field one() {
        Field r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = P::one(i);
        return r;
    }
    BMPC_HD static Field r2() {
        Field r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = P::r2(i);
        return r;
    }
    BMPC_HD bool is_zero() const {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < N; i++) acc |= l[i];
        return acc == 0;
    }
    BMPC_HD bool operator==(const Field& o) const {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < N; i++) acc |= l[i] ^ o.l[i];
        return acc == 0;
    }
    BMPC_HD bool operator!=(const Field& o) const { return !(*this == o); }

    // r = a - p if a >= p else a   (a < 2p)
    BMPC_HD static void final_sub(uint32_t* r, const uint32_t* a) {
        uint32_t t[N];
        t[0] = sub_cc(a[0], P::mod(0));
#pragma unroll
        for (int i = 1; i < N; i++) t[i] = subc_cc(a[i], P::mod(i));
        uint32_t borrow = subc(0, 0);  // 0 or 0xffffffff
#pragma unroll
        for (int i = 0; i < N; i++) r[i] = borrow ? a[i] : t[i];
    }

    BMPC_HD friend Field operator+(const Field& a, const Field& b) {
        uint32_t t[N];
        t[0] = add_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N; i++) t[i] = addc_cc(a.l[i], b.l[i]);
        // both moduli leave the top bit(s) of the top limb free: no carry out of limb N-1
        Field r;
        final_sub(r.l, t);
        return r;
    }
    BMPC_HD friend Field operator-(const Field& a, const Field& b) {
        uint32_t t[N];
        t[0] = sub_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N; i++) t[i] = subc_cc(a.l[i], b.l[i]);
        uint32_t borrow = subc(0, 0);
        Field r;
        r.l[0] = add_cc(t[0], borrow & P::mod(0));
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(t[i], borrow & P::mod(i));
        r.l[N - 1] = addc(t[N - 1], borrow & P::mod(N - 1));
        return r;
    }
    BMPC_HD Field neg() const {
        if (is_zero()) return *this;
        Field r;
        r.l[0] = sub_cc(P::mod(0), l[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = subc_cc(P::mod(i), l[i]);
        r.l[N - 1] = subc(P::mod(N - 1), l[N - 1]);
        return r;
    }
    BMPC_HD Field dbl() const { return *this + *this; }

    // One CIOS step: (even, odd) <- ((even, odd) >> 32) + a * bi, then add mi * p so that
    // limb 0 vanishes.  `even` holds the 2-limb products whose low limb sits at an even
    // position, `odd` those at an odd position (stored shifted down by one limb).
    BMPC_HD static void mad_redc(uint32_t* even, uint32_t* odd, const uint32_t* a, uint32_t bi,
                                 bool first) {
        if (first) {
#pragma unroll
            for (int j = 0; j < N; j += 2) {
                odd[j] = mul_lo(a[j + 1], bi);
                odd[j + 1] = mul_hi(a[j + 1], bi);
                even[j] = mul_lo(a[j], bi);
                even[j + 1] = mul_hi(a[j], bi);
            }
        } else {
            // `odd` is last step's even array: its limb 0 is zero, its limb 1 lands on
            // this step's limb 0, limbs 2.. land on odd positions again (shift by two).
            even[0] = add_cc(even[0], odd[1]);
#pragma unroll
            for (int j = 0; j < N - 2; j += 2) {
                odd[j] = madc_lo_cc(a[j + 1], bi, odd[j + 2]);
                odd[j + 1] = madc_hi_cc(a[j + 1], bi, odd[j + 3]);
            }
            odd[N - 2] = madc_lo_cc(a[N - 1], bi, 0);
            odd[N - 1] = madc_hi(a[N - 1], bi, 0);
            even[0] = mad_lo_cc(a[0], bi, even[0]);
            even[1] = madc_hi_cc(a[0], bi, even[1]);
#pragma unroll
            for (int j = 2; j < N; j += 2) {
                even[j] = madc_lo_cc(a[j], bi, even[j]);
                even[j + 1] = madc_hi_cc(a[j], bi, even[j + 1]);
            }
            odd[N - 1] = addc(odd[N - 1], 0);
        }
        uint32_t mi = even[0] * P::INV;
        odd[0] = mad_lo_cc(P::mod(1), mi, odd[0]);
        odd[1] = madc_hi_cc(P::mod(1), mi, odd[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            odd[j] = madc_lo_cc(P::mod(j + 1), mi, odd[j]);
            odd[j + 1] = madc_hi_cc(P::mod(j + 1), mi, odd[j + 1]);
        }
        even[0] = mad_lo_cc(P::mod(0), mi, even[0]);
        even[1] = madc_hi_cc(P::mod(0), mi, even[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            even[j] = madc_lo_cc(P::mod(j), mi, even[j]);
            even[j + 1] = madc_hi_cc(P::mod(j), mi, even[j + 1]);
        }
        odd[N - 1] = addc(odd[N - 1], 0);
    }

    // Montgomery product a * b * R^-1 mod p, fully reduced.
    BMPC_HD friend Field operator*(const Field& a, const Field& b) {
        uint32_t even[N], odd[N];
        // Both operands are read into registers BEFORE the first carry-chain statement.  When a or b lives
        // in memory (shared-memory butterflies, twiddle tables, by-reference arguments) the compiler will not
        // merge loads across the volatile asm statements: the lo and the hi half of a limb product then
        // multiply two different registers and ptxas cannot pair them into one IMAD.WIDE.X -- the product
        // becomes IMAD + IMAD.HI + two IADD3 per limb pair (Fr: 64 + 64 + 65 multiplier instructions instead
        // of 129; the whole NTT and every out-of-line curve operation ran that way until round 3).
        uint32_t al[N], bl[N];
#pragma unroll
        for (int j = 0; j < N; j++) { al[j] = a.l[j]; bl[j] = b.l[j]; }
#pragma unroll
        for (int i = 0; i < N; i += 2) {
            mad_redc(even, odd, al, bl[i], i == 0);
            mad_redc(odd, even, al, bl[i + 1], false);
        }
        // last step left the reduced limb in even[0] (== 0); merge the two chains
        even[0] = add_cc(even[0], odd[1]);
#pragma unroll
        for (int j = 1; j < N - 1; j++) even[j] = addc_cc(even[j], odd[j + 1]);
        even[N - 1] = addc(even[N - 1], 0);
        Field r;
        final_sub(r.l, even);
        return r;
    }
    // Montgomery square a * a * R^-1 mod p: N(N-1)/2 off-diagonal products (doubled) + N
    // diagonal ones, then a word-by-word REDC -- (N^2+N)/2 + N^2 + N MAC32 instead of 2N^2 + N
    // (234 vs 300 for Fp).
    // Measured on B200 inside the MSM accumulate kernel this is NOT faster than operator* (7 %
    // fewer IMADs but ~900 extra register moves and 30 more registers), so sqr() below still
    // multiplies; kept, and host-tested, for the next tuning round.
    BMPC_HD Field sqr_redc() const {
        constexpr int N2 = 2 * N;
        const uint32_t* a = l;
        uint32_t d[N2];
#pragma unroll
        for (int k = 0; k < N2; k++) d[k] = 0;
        // Row i adds a_i * a_j (j > i) at limb i + j.  Products whose offset k = j - i has the
        // same parity tile contiguous limb pairs, so each parity class is one carry chain.  The
        // class NOT holding the last term ends one limb lower and runs first: its carry-out then
        // lands on a limb that so far only holds a carry bit, and the second chain's carry-out on
        // an untouched limb (rows grow the top by one limb each), so no ripple is ever needed.
#pragma unroll
        for (int i = 0; i < N - 1; i++) {
            const int K = N - 1 - i;
#pragma unroll
            for (int pass = 0; pass < 2; pass++) {
                const int ks = (pass == 0) ? ((K & 1) ? 2 : 1) : ((K & 1) ? 1 : 2);
                if (ks <= K) {
                    int last = 0;
#pragma unroll
                    for (int k = ks; k <= K; k += 2) {
                        const int j = i + k, pos = i + j;
                        if (k == ks) d[pos] = mad_lo_cc(a[i], a[j], d[pos]);
                        else d[pos] = madc_lo_cc(a[i], a[j], d[pos]);
                        d[pos + 1] = madc_hi_cc(a[i], a[j], d[pos + 1]);
                        last = pos + 1;
                    }
                    d[last + 1] = addc(d[last + 1], 0);
                }
            }
        }
        // t = 2 d + sum a_i^2 2^(64 i)
        uint32_t t[N2];
#pragma unroll
        for (int k = N2 - 1; k >= 1; k--) t[k] = (d[k] << 1) | (d[k - 1] >> 31);
        t[0] = 0;
#pragma unroll
        for (int i = 0; i < N; i++) {
            if (i == 0) t[0] = mad_lo_cc(a[0], a[0], t[0]);
            else t[2 * i] = madc_lo_cc(a[i], a[i], t[2 * i]);
            t[2 * i + 1] = madc_hi_cc(a[i], a[i], t[2 * i + 1]);
        }
        // REDC: for each low limb add m p so it vanishes.  The two carry-outs of row i land on
        // limbs i+N and i+N+1; they are counted separately and folded in once at the end, so
        // the rows need no ripple either.
        uint32_t cy[N + 1];
#pragma unroll
        for (int k = 0; k <= N; k++) cy[k] = 0;
#pragma unroll
        for (int i = 0; i < N; i++) {
            uint32_t m = t[i] * P::INV;
            t[i] = mad_lo_cc(m, P::mod(0), t[i]);
            t[i + 1] = madc_hi_cc(m, P::mod(0), t[i + 1]);
#pragma unroll
            for (int j = 2; j < N; j += 2) {
                t[i + j] = madc_lo_cc(m, P::mod(j), t[i + j]);
                t[i + j + 1] = madc_hi_cc(m, P::mod(j), t[i + j + 1]);
            }
            cy[i] = addc(cy[i], 0);
            t[i + 1] = mad_lo_cc(m, P::mod(1), t[i + 1]);
            t[i + 2] = madc_hi_cc(m, P::mod(1), t[i + 2]);
#pragma unroll
            for (int j = 3; j < N; j += 2) {
                t[i + j] = madc_lo_cc(m, P::mod(j), t[i + j]);
                t[i + j + 1] = madc_hi_cc(m, P::mod(j), t[i + j + 1]);
            }
            cy[i + 1] = addc(cy[i + 1], 0);
        }
        uint32_t r[N];
        r[0] = add_cc(t[N], cy[0]);
#pragma unroll
        for (int k = 1; k < N - 1; k++) r[k] = addc_cc(t[N + k], cy[k]);
        r[N - 1] = addc(t[2 * N - 1], cy[N - 1]);
        Field out;
        final_sub(out.l, r);
        return out;
    }
    BMPC_HD Field sqr() const { return *this * *this; }
    // out-of-line product for cold paths
    // By VALUE: the operands travel in registers and the body is the same straight IMAD.WIDE.X run as the
    // inlined product.  Taking them by reference made the compiler load limbs from the caller's stack in the
    // middle of the carry chains, which broke the lo/hi pairing: 134 IMAD.WIDE + 156 IMAD + 156 IMAD.HI + 322
    // IADD3 per product instead of 278 IMAD.WIDE + 57 IADD3 (every kernel of the bucket reduction, the quad
    // steps, the table precomputation and the G2 product tree ran on that body until round 3).
    BMPC_COLD static Field mul_cold(Field a, Field b) { return a * b; }

    BMPC_HD Field to_mont() const { return *this * r2(); }
    BMPC_HD Field from_mont() const {
        Field o = zero();
        o.l[0] = 1;
        return *this * o;
    }

    // this^e for a little-endian multi-word exponent (vartime; setup / one-off use only)
    BMPC_COLD Field pow(const uint32_t* e, int words) const {
        Field r = one();
        bool started = false;
        for (int i = words - 1; i >= 0; i--) {
            for (int b = 31; b >= 0; b--) {
                if (started) r = mul_cold(r, r);
                if ((e[i] >> b) & 1) {
                    r = mul_cold(r, *this);
                    started = true;
                }
            }
        }
        return r;
    }
    BMPC_HD Field pow_u64(uint64_t e) const {
        uint32_t w[2] = {(uint32_t)e, (uint32_t)(e >> 32)};
        return pow(w, 2);
    }
    // Fermat inverse (0 -> 0): ~1.5 log p products; kept as the reference for inv().
    BMPC_COLD Field inv_fermat() const {
        uint32_t e[N];
        e[0] = sub_cc(P::mod(0), 2);
#pragma unroll
        for (int i = 1; i < N; i++) e[i] = subc_cc(P::mod(i), 0);
        return pow(e, N);
    }

    // x <- x / 2 mod p   (x < p)
    BMPC_HD void halve() {
        uint32_t odd = l[0] & 1u;
        uint32_t t[N];
        t[0] = add_cc(l[0], odd ? P::mod(0) : 0u);
#pragma unroll
        for (int i = 1; i < N; i++) t[i] = addc_cc(l[i], odd ? P::mod(i) : 0u);
        // both moduli leave the top bit free, so x + p does not overflow N limbs
#pragma unroll
        for (int i = 0; i < N - 1; i++) l[i] = (t[i] >> 1) | (t[i + 1] << 31);
        l[N - 1] = t[N - 1] >> 1;
    }

    // Inverse (0 -> 0).  It sits on the serial path of every block-shared batch inversion of the bucket
    // accumulation (msm_affine.cuh: one thread inverts while the block waits; measured with the
    // inversion stubbed out: 12.8 % of the accumulate kernel at 2^21 points, 3.9 % at 2^24) and of every
    // to_affine, so instructions, not products, are what counts.  Kaliski's almost-inverse (1995): the
    // binary extended Euclid without any modular reduction inside the loop --
    //     u = p, v = a, r = 0, s = 1;   invariants  a r = -u 2^k,  a s = v 2^k (mod p),  u s + v r = p
    //     u even: u /= 2, s *= 2 | v even: v /= 2, r *= 2 | u > v: u = (u-v)/2, r += s, s *= 2
    //     | else: v = (v-u)/2, s += r, r *= 2;   k counts the halvings
    // -- with every run of trailing zero bits taken in one multi-limb shift.  It ends with
    // p - r = a^-1 2^k, bits(p) <= k <= 2 bits(p); the power of two goes away in two Montgomery products:
    // (x R^2 / R) (2^(2 log R - k)) / R = x R^2 / 2^k.  About a third of the instructions of the textbook
    // binary algorithm it replaces (which halved x1, x2 modulo p at every step).  Vartime, like the
    // reference's `invert` callers here.
    BMPC_HD static uint32_t ctz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
        return (uint32_t)(__ffs((int)x) - 1);
#else
        return (uint32_t)__builtin_ctz(x);
#endif
    }
    // x >>= t, 1 <= t <= 31
    BMPC_HD static void shr_limbs(uint32_t* x, uint32_t t) {
#pragma unroll
        for (int i = 0; i < N - 1; i++) x[i] = (x[i] >> t) | (x[i + 1] << (32u - t));
        x[N - 1] >>= t;
    }
    // x <<= t, 1 <= t <= 31 (the caller guarantees no bit is lost)
    BMPC_HD static void shl_limbs(uint32_t* x, uint32_t t) {
#pragma unroll
        for (int i = N - 1; i > 0; i--) x[i] = (x[i] << t) | (x[i - 1] >> (32u - t));
        x[0] <<= t;
    }
    BMPC_COLD Field inv() const {
        if (is_zero()) return *this;
        uint32_t u[N], v[N], r[N], s[N];
#pragma unroll
        for (int i = 0; i < N; i++) { u[i] = P::mod(i); v[i] = l[i]; r[i] = 0; s[i] = 0; }
        s[0] = 1;
        uint32_t k = 0;
        for (;;) {
            if (!(u[0] & 1u)) {                    // u > 0 always: a zero low limb is a run of >= 32 zeros
                const uint32_t t = u[0] ? ctz32(u[0]) : 31u;
                shr_limbs(u, t);
                shl_limbs(s, t);
                k += t;
                continue;
            }
            if (!(v[0] & 1u)) {
                const uint32_t t = v[0] ? ctz32(v[0]) : 31u;
                shr_limbs(v, t);
                shl_limbs(r, t);
                k += t;
                continue;
            }
            uint32_t d[N];
            d[0] = sub_cc(u[0], v[0]);
#pragma unroll
            for (int i = 1; i < N; i++) d[i] = subc_cc(u[i], v[i]);
            const uint32_t borrow = subc(0, 0);
            uint32_t nz = 0;
#pragma unroll
            for (int i = 0; i < N; i++) nz |= d[i];
            if (!borrow && nz) {                   // u > v
#pragma unroll
                for (int i = 0; i < N; i++) u[i] = d[i];
                shr_limbs(u, 1);
                r[0] = add_cc(r[0], s[0]);
#pragma unroll
                for (int i = 1; i < N - 1; i++) r[i] = addc_cc(r[i], s[i]);
                r[N - 1] = addc(r[N - 1], s[N - 1]);
                shl_limbs(s, 1);
                k++;
            } else {                                // v >= u
                v[0] = sub_cc(v[0], u[0]);
#pragma unroll
                for (int i = 1; i < N - 1; i++) v[i] = subc_cc(v[i], u[i]);
                v[N - 1] = subc(v[N - 1], u[N - 1]);
                shr_limbs(v, 1);
                s[0] = add_cc(s[0], r[0]);
#pragma unroll
                for (int i = 1; i < N - 1; i++) s[i] = addc_cc(s[i], r[i]);
                s[N - 1] = addc(s[N - 1], r[N - 1]);
                shl_limbs(r, 1);
                k++;
                if (!nz) break;                     // u == v (== 1, the gcd): v is now 0
            }
        }
        // r < 2 p;  x = p - (r mod p) = a^-1 2^k as a raw integer (a = the raw limbs of *this)
        Field x;
        final_sub(x.l, r);
        x = x.neg();
        // wanted: (a / R)^-1 R = a^-1 R^2 = x 2^e with e = 2 log2 R - k; the single-limb constant 2^e must
        // stay below p, i.e. e <= bits(p) - 1 (k within a few bits of its minimum otherwise: double x)
        uint32_t e = 64u * N - k;
        while (e >= 32u * N - lead_zero_bits()) { x = x.dbl(); e--; }
        Field c = zero();
        c.l[e >> 5] = 1u << (e & 31u);
        return mul_cold(mul_cold(x, r2()), c);      // (x R^2 / R) 2^e / R = x 2^e
    }
    // number of zero bits above the modulus' top bit in its top limb (Fp: 3, Fr: 1)
    BMPC_HD static constexpr uint32_t lead_zero_bits() {
        uint32_t t = P::mod(N - 1), z = 0;
        while (!(t & 0x80000000u)) { t <<= 1; z++; }
        return z;
    }
};

typedef Field<FrParams> Fr;
typedef Field<FpParams> Fp;

// ------------------------------------------------------------------------------- Fp2
// Fp2 = Fp[u]/(u^2 + 1)  (bls12_381::Fp2; third-party).
struct Fp2 {
    Fp c0, c1;
    BMPC_HD static Fp2 zero() { return Fp2{Fp::zero(), Fp::zero()}; }
    BMPC_HD static Fp2 one() { return Fp2{Fp::one(), Fp::zero()}; }
    BMPC_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    BMPC_HD bool operator==(const Fp2& o) const { return c0 == o.c0 && c1 == o.c1; }
    BMPC_HD bool operator!=(const Fp2& o) const { return !(*this == o); }
    BMPC_HD friend Fp2 operator+(const Fp2& a, const Fp2& b) { return Fp2{a.c0 + b.c0, a.c1 + b.c1}; }
    BMPC_HD friend Fp2 operator-(const Fp2& a, const Fp2& b) { return Fp2{a.c0 - b.c0, a.c1 - b.c1}; }
    BMPC_HD Fp2 neg() const { return Fp2{c0.neg(), c1.neg()}; }
    BMPC_HD Fp2 dbl() const { return Fp2{c0.dbl(), c1.dbl()}; }
    // BMPC_FP2_CALLS (device): the three Fp products of every Fp2 product are calls of the out-of-line
    // routine (operands by value, in registers) instead of three inlined bodies.
#if defined(__CUDA_ARCH__) && defined(BMPC_FP2_CALLS)
    BMPC_HD friend Fp2 operator*(const Fp2& a, const Fp2& b) {
        Fp t0 = Fp::mul_cold(a.c0, b.c0);
        Fp t1 = Fp::mul_cold(a.c1, b.c1);
        Fp t2 = Fp::mul_cold(a.c0 + a.c1, b.c0 + b.c1);
        return Fp2{t0 - t1, t2 - t0 - t1};
    }
    BMPC_HD Fp2 sqr() const {
        Fp s = c0 + c1;
        Fp d = c0 - c1;
        Fp m = Fp::mul_cold(c0, c1);
        return Fp2{Fp::mul_cold(s, d), m.dbl()};
    }
#else
    BMPC_HD friend Fp2 operator*(const Fp2& a, const Fp2& b) {
        Fp t0 = a.c0 * b.c0;
        Fp t1 = a.c1 * b.c1;
        Fp t2 = (a.c0 + a.c1) * (b.c0 + b.c1);
        return Fp2{t0 - t1, t2 - t0 - t1};
    }
    BMPC_HD Fp2 sqr() const {
        Fp s = c0 + c1;
        Fp d = c0 - c1;
        Fp m = c0 * c1;
        return Fp2{s * d, m.dbl()};
    }
#endif
    BMPC_COLD Fp2 inv() const {
        Fp n = (Fp::mul_cold(c0, c0) + Fp::mul_cold(c1, c1)).inv();
        return Fp2{Fp::mul_cold(c0, n), Fp::mul_cold(c1, n).neg()};
    }
    // compact cold product: three calls of the out-of-line Fp product (keeps cold kernels small
    // enough for the instruction cache -- see curve.cuh)
    BMPC_COLD static Fp2 mul_cold(const Fp2& a, const Fp2& b) {
        Fp t0 = Fp::mul_cold(a.c0, b.c0);
        Fp t1 = Fp::mul_cold(a.c1, b.c1);
        Fp t2 = Fp::mul_cold(a.c0 + a.c1, b.c0 + b.c1);
        return Fp2{t0 - t1, t2 - t0 - t1};
    }
};

}  // namespace bmpc
