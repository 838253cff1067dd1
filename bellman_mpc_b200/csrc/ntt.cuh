// Radix-2 Fr number-theoretic transform for EvaluationDomain::{fft, ifft, coset_fft,
// icoset_fft} (reference: src/domain.rs:81-125, best_fft/serial_fft/parallel_fft :261-372).
//
// The reference bit-reverses and runs log n DIT stages over the whole vector (serial_fft) or
// splits into 2^log_cpus sub-FFTs (parallel_fft).  Here the transform is a Stockham autosort
// decomposition n = R_1 * R_2 * ... with R_j = 2^deg_j <= 2^10: one kernel launch per factor,
// each block running all deg_j butterfly stages of T independent R_j-point sub-transforms out
// of shared memory, so a 2^24 transform touches HBM 3 times instead of 24.  Natural order in,
// natural order out (SURVEY 8a'/9) -- no separate bit-reversal pass.  The coset shift g^i,
// the 1/m of the inverse transform and (for the prover) 1/Z are folded into the first pass's
// loads / the last pass's stores (distribute_powers :101-113, ifft's minv loop :88-98,
// divide_by_z_on_coset :139-151) instead of being separate sweeps over memory.
//
// Algorithmic traffic per transform: 64 B per coefficient (32 B read + 32 B write);
// (m/2) log2 m butterflies of one Fr Montgomery product (136 MAC32) each.
#pragma once
#include "field.cuh"

namespace bmpc {

// x^e = hi[e >> lo_bits] * lo[e & mask]; a constant may be folded into every `lo` entry.
struct PowTable {
    const Fr* hi;
    const Fr* lo;
    uint32_t lo_bits;
    uint32_t hi_n;      // entries in hi; 1 => lo alone covers the range
    const Fr* direct;   // optional: all 2^logm powers expanded (one load, no product); may be NULL
};

enum { SCALE_NONE = 0, SCALE_CONST = 1, SCALE_POW = 2 };

struct NttPassArgs {
    const Fr* in;
    Fr* out;
    uint32_t logn, deg, plog, tile_log;
    PowTable tw;         // powers of omega_n (or its inverse)
    const Fr* tw_small;  // powers of the 2^small_log-th root: tw_small[j], j < 2^(small_log-1)
    uint32_t small_log;
    int pre_mode, post_mode;
    PowTable pre, post;
    Fr post_const;
    size_t batch_stride;  // elements between consecutive transforms of a batch (gridDim.y of them)
};

__device__ __forceinline__ Fr load_fr(const Fr* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ Fr ldg_fr(const Fr* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void store_fr(Fr* p, const Fr& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
__device__ __forceinline__ Fr pow_lookup(const PowTable& t, uint32_t e) {
    if (t.direct) return ldg_fr(t.direct + e);
    Fr lo = ldg_fr(t.lo + (e & ((1u << t.lo_bits) - 1u)));
    if (t.hi_n == 1) return lo;
    return lo * ldg_fr(t.hi + (e >> t.lo_bits));
}

// One Stockham pass.  With t = n / R, p = 2^plog (product of the radices already done):
//   for i in [0, t), k = i mod p:
//     v[s]  = in[i + s t] * omega_n^{(t/p) k s}                 s in [0, R)
//     V     = DFT_R(v)            (DIF stages in shared memory, read out bit-reversed)
//     out[(i - k) R + k + s' p] = V[s']
// Block = T consecutive i, R = 2^deg up to 2^12 (128 KB of shared memory: a 2^24 transform is two
// passes); threads loop over the T R / 2 butterflies of a stage.  Shared layout u[s][i_local].
__global__ void __launch_bounds__(1024) ntt_pass_kernel(NttPassArgs A) {
    extern __shared__ uint4 ntt_smem[];
    Fr* u = reinterpret_cast<Fr*>(ntt_smem);
    const uint32_t deg = A.deg, R = 1u << deg, T = 1u << A.tile_log;
    const uint32_t tlog = A.logn - deg;
    const uint32_t tid = threadIdx.x, nthr = blockDim.x;
    const uint32_t tile_base = blockIdx.x << A.tile_log;
    const uint32_t pmask = (1u << A.plog) - 1u;
    const uint32_t total = T << deg;              // elements held by the block
    const Fr* in = A.in + (size_t)blockIdx.y * A.batch_stride;
    Fr* out = A.out + (size_t)blockIdx.y * A.batch_stride;

    for (uint32_t e = tid; e < total; e += nthr) {
        uint32_t il = e & (T - 1u), s = e >> A.tile_log;
        uint32_t i = tile_base + il;
        uint32_t src = i + (s << tlog);
        Fr v = load_fr(in + src);
        if (A.pre_mode == SCALE_POW) v = v * pow_lookup(A.pre, src);
        if (A.plog != 0) {
            uint32_t ex = ((i & pmask) * s) << (tlog - A.plog);
            if (ex != 0) v = v * pow_lookup(A.tw, ex);
        }
        u[e] = v;                                  // u[s * T + il]
    }
    __syncthreads();

    const uint32_t sshift = A.small_log - deg;
#pragma unroll 1
    for (uint32_t rnd = 0; rnd < deg; rnd++) {
        uint32_t half = R >> (rnd + 1);
        for (uint32_t bb = tid; bb < (total >> 1); bb += nthr) {
            uint32_t il = bb & (T - 1u), b = bb >> A.tile_log;
            uint32_t di = b & (half - 1u);
            uint32_t i0 = ((b - di) << 1) + di;
            uint32_t i1 = i0 + half;
            Fr x0 = u[i0 * T + il];
            Fr x1 = u[i1 * T + il];
            u[i0 * T + il] = x0 + x1;
            Fr d = x0 - x1;
            if (di != 0) d = d * ldg_fr(A.tw_small + ((di << rnd) << sshift));
            u[i1 * T + il] = d;
        }
        __syncthreads();
    }

    for (uint32_t e = tid; e < total; e += nthr) {
        uint32_t sp, il2;
        if (A.plog >= A.tile_log) {  // consecutive threads -> consecutive k: contiguous stores
            il2 = e & (T - 1u);
            sp = e >> A.tile_log;
        } else {                      // first pass (p < T): consecutive threads -> consecutive s'
            sp = e & (R - 1u);
            il2 = e >> deg;
        }
        uint32_t i2 = tile_base + il2;
        uint32_t k2 = i2 & pmask;
        uint32_t dst = ((i2 - k2) << deg) + k2 + (sp << A.plog);
        Fr v = u[(__brev(sp) >> (32u - deg)) * T + il2];
        if (A.post_mode == SCALE_CONST) v = v * A.post_const;
        else if (A.post_mode == SCALE_POW) v = v * pow_lookup(A.post, dst);
        store_fr(out + dst, v);
    }
}

// ------------------------------------------------- pieces of the distributed four-step transform
// out[b][a][c] = in[a][b][c] for 32-byte elements, d2 >= 1 consecutive elements moved together
// (d2 == 1 is a plain d0 x d1 -> d1 x d0 transpose).  32 x 32 tiles of (a, b) go through shared
// memory so that both the loads (consecutive b) and the stores (consecutive a) are contiguous runs
// when d2 is small; for d2 >= 32 the c index alone gives coalescing and the tile is 1 x 1 logically.
#define BMPC_TR_TILE 16
__global__ void __launch_bounds__(256) fr_swap01_kernel(const Fr* in, Fr* out, uint32_t d0, uint32_t d1, uint32_t d2) {
    __shared__ uint4 tile[BMPC_TR_TILE][BMPC_TR_TILE + 1][2];
    const uint32_t tx = threadIdx.x & (BMPC_TR_TILE - 1), ty = threadIdx.x / BMPC_TR_TILE;
    const uint32_t a0 = blockIdx.y * BMPC_TR_TILE, b0 = blockIdx.x * BMPC_TR_TILE;
    for (uint32_t c = blockIdx.z; c < d2; c += gridDim.z) {
        uint32_t a = a0 + ty, b = b0 + tx;
        if (a < d0 && b < d1) {
            const uint4* q = reinterpret_cast<const uint4*>(in + ((size_t)a * d1 + b) * d2 + c);
            tile[ty][tx][0] = q[0];
            tile[ty][tx][1] = q[1];
        }
        __syncthreads();
        a = a0 + tx; b = b0 + ty;
        if (a < d0 && b < d1) {
            uint4* q = reinterpret_cast<uint4*>(out + ((size_t)b * d0 + a) * d2 + c);
            q[0] = tile[tx][ty][0];
            q[1] = tile[tx][ty][1];
        }
        __syncthreads();
    }
}
// same permutation when the inner run is long: one thread per element, consecutive threads ->
// consecutive c (coalesced on both sides)
__global__ void fr_swap01_rows_kernel(const Fr* in, Fr* out, uint32_t d0, uint32_t d1, uint32_t d2) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)d0 * d1 * d2;
    if (idx >= total) return;
    uint32_t c = (uint32_t)(idx % d2);
    size_t ab = idx / d2;
    uint32_t b = (uint32_t)(ab % d1), a = (uint32_t)(ab / d1);
    store_fr(out + ((size_t)b * d0 + a) * d2 + c, load_fr(in + idx));
}
// d[r][c] *= w^((row0 + r) c), w = omega_m or its inverse (`tw`: powers of w, m = 2^logm):
// the twiddle step between the column and the row transforms of the four-step decomposition.
__global__ void fr_fourstep_twiddle_kernel(Fr* d, uint32_t rows, uint32_t cols, uint32_t row0, uint32_t logm,
                                           PowTable tw) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)rows * cols) return;
    uint32_t c = (uint32_t)(idx % cols), r = (uint32_t)(idx / cols);
    uint64_t e = ((uint64_t)(row0 + r) * c) & (((uint64_t)1 << logm) - 1u);
    if (e) store_fr(d + idx, load_fr(d + idx) * pow_lookup(tw, (uint32_t)e));
}
// d[i] *= T[first + i] (coset shift g^i, or g^-i / m), or d[i] *= *k when T.lo == NULL
__global__ void fr_scale_pow_kernel(Fr* d, size_t n, uint32_t first, PowTable T, const Fr* k) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr f = T.lo ? pow_lookup(T, first + (uint32_t)i) : *k;
    store_fr(d + i, load_fr(d + i) * f);
}

// direct[e] = hi[e >> lo_bits] * lo[e & mask]  (expands a two-level table once)
__global__ void pow_expand_kernel(PowTable t, uint32_t count, Fr* out) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count) return;
    Fr lo = ldg_fr(t.lo + (e & ((1u << t.lo_bits) - 1u)));
    store_fr(out + e, t.hi_n == 1 ? lo : lo * ldg_fr(t.hi + (e >> t.lo_bits)));
}

// n == 1 transform: only the scalings apply.
__global__ void ntt_trivial_kernel(NttPassArgs A) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        Fr v = load_fr(A.in);
        if (A.pre_mode == SCALE_POW) v = v * pow_lookup(A.pre, 0);
        if (A.post_mode == SCALE_CONST) v = v * A.post_const;
        else if (A.post_mode == SCALE_POW) v = v * pow_lookup(A.post, 0);
        store_fr(A.out, v);
    }
}

// ------------------------------------------------------------------ setup kernels (tiny)
// Domain constants for m = 2^logm (src/domain.rs:62-77,129-151), all Montgomery:
//  [0] omega  [1] omega^-1  [2] m^-1  [3] g = 7  [4] g^-1  [5] (g^m - 1)^-1
//  [6] m^-1 * (g^m - 1)^-1  [7] one  [8] omega_small  [9] omega_small^-1
// root_of_unity = 7^((q-1)/2^32) (ff::PrimeField::root_of_unity for bls12_381::Scalar).
__global__ void domain_consts_kernel(uint32_t logm, uint32_t small_log, Fr* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Fr root;  // Montgomery limbs of root_of_unity (oracle/fields.py self_check pins the value)
    const uint32_t rl[8] = {0x5f0e466au, 0xb9b58d8cu, 0x1819d7ecu, 0x5b1b4c80u,
                            0x52a31e64u, 0x0af53ae3u, 0x19e9b27bu, 0x5bf3addau};
    for (int j = 0; j < 8; j++) root.l[j] = rl[j];
    Fr omega = root;
    for (uint32_t j = logm; j < 32; j++) omega = omega.sqr();
    Fr osmall = root;
    for (uint32_t j = small_log; j < 32; j++) osmall = osmall.sqr();
    Fr one = Fr::one();
    Fr seven = one;
    for (int j = 0; j < 6; j++) seven = seven + one;
    Fr m = one;
    for (uint32_t j = 0; j < logm; j++) m = m.dbl();
    Fr gm = seven;
    for (uint32_t j = 0; j < logm; j++) gm = gm.sqr();
    Fr zinv = (gm - one).inv();
    Fr minv = m.inv();
    out[0] = omega;
    out[1] = omega.inv();
    out[2] = minv;
    out[3] = seven;
    out[4] = seven.inv();
    out[5] = zinv;
    out[6] = minv * zinv;
    out[7] = one;
    out[8] = osmall;
    out[9] = osmall.inv();
}

// out[j] = fold * base^(j * stride), j < count   (fold == NULL -> 1).  With canon != 0 the
// entry is stored in canonical (non-Montgomery) form, so that a Montgomery product with it
// both scales a coefficient and strips its Montgomery factor (to_le_bits for free).
__global__ void pow_table_kernel(const Fr* base, const Fr* fold, uint64_t stride, uint32_t count,
                                 int canon, Fr* out) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    Fr v = base->pow_u64((uint64_t)j * stride);
    if (fold) v = v * *fold;
    if (canon) v = v.from_mont();
    out[j] = v;
}
// out[j] = a[j] * b[j] (tiny, for folding two constants)
__global__ void fr_mul_kernel(const Fr* a, const Fr* b, Fr* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = *a * *b;
}

// ------------------------------------------------------------------- pointwise kernels
// mul_assign (:154-170), sub_assign (:173-189), constant scaling (divide_by_z :139-151)
__global__ void fr_mul_assign_kernel(Fr* a, const Fr* b, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) store_fr(a + i, load_fr(a + i) * load_fr(b + i));
}
__global__ void fr_sub_assign_kernel(Fr* a, const Fr* b, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) store_fr(a + i, load_fr(a + i) - load_fr(b + i));
}
__global__ void fr_scale_kernel(Fr* a, const Fr* k, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    Fr kk = *k;
    if (i < n) store_fr(a + i, load_fr(a + i) * kk);
}
// a <- a * b - c  (prover.rs:221-224 mul_assign + sub_assign fused; 1/Z is folded into the
// following icoset pass)
__global__ void fr_mul_sub_kernel(Fr* a, const Fr* b, const Fr* c, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) store_fr(a + i, load_fr(a + i) * load_fr(b + i) - load_fr(c + i));
}
// distribute_powers(g) for an arbitrary g (:101-113): each thread owns CH consecutive
// coefficients, starts from g^(first index) and keeps a running product like the reference's
// per-chunk loop.
__global__ void fr_distribute_powers_kernel(Fr* a, const Fr* g, size_t n) {
    const int CH = 8;
    size_t first = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * CH;
    if (first >= n) return;
    Fr gg = *g;
    Fr u = gg.pow_u64(first);
    for (int j = 0; j < CH && first + j < n; j++) {
        store_fr(a + first + j, load_fr(a + first + j) * u);
        u = u * gg;
    }
}
// Montgomery -> canonical little-endian (PrimeFieldBits::to_le_bits, prover.rs:231,241,248)
__global__ void fr_to_canonical_kernel(Fr* a, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) store_fr(a + i, load_fr(a + i).from_mont());
}
__global__ void fr_zero_kernel(Fr* a, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) store_fr(a + i, Fr::zero());
}
// tau^m - 1 (domain.rs:129-134)
__global__ void fr_z_kernel(const Fr* tau, uint32_t logm, Fr* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Fr t = *tau;
    for (uint32_t j = 0; j < logm; j++) t = t.sqr();
    *out = t - Fr::one();
}

}  // namespace bmpc
