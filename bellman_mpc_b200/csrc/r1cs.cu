// Constraint-system side of the prover and the key generator, on the
// device, over R1CS matrices in CSR form:
//
//  * bmpc_r1cs_eval          = ProvingAssignment::enforce for the whole system at once
//                              (src/groth16/prover.rs:19-53,100-138) + the `x * 0 = 0` input rows of
//                              create_proof (:202-204): a, b, c evaluations and the three density maps
//  * bmpc_generate_parameters = generate_parameters with upstream semantics
//                              (src/groth16/generator.rs:241-272,294-297,310-572,584-590,594-604,612-634;
//                              the fork's MPC cross-check hooks :273-292,298-308,573-577,592-593,605-611
//                              only assert/print and break beyond 4-constraint toys -- SURVEY 4)
//
// Both are sparse matrix x vector products over Fr (one Montgomery product per non-zero) around the
// kernels that already exist (NTT, fixed-base multiplication); nothing here runs field arithmetic
// on the host.
#include <cstring>
#include <vector>

#include "internal.h"

namespace bmpc {

namespace {
__device__ __forceinline__ Fr load_fr(const Fr* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
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
}  // namespace

// out[i] = tau^i (generator.rs:351-366)
__global__ void tau_powers_kernel(const Fr* tau, uint32_t count, Fr* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) store_fr(out + i, tau->pow_u64(i));
}

struct DevCsr {
    const uint32_t* row_ptr;
    const uint32_t* col;
    const Fr* coeff;
    uint32_t num_rows;
};

// y[r] = sum_k coeff[k] * x[col[k]]  (+ extra[r] if extra != NULL); one thread per row
__global__ void spmv_thread_kernel(DevCsr m, const Fr* x, const Fr* extra, Fr* y) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m.num_rows) return;
    Fr acc = extra ? load_fr(extra + r) : Fr::zero();
    for (uint32_t k = m.row_ptr[r]; k < m.row_ptr[r + 1]; k++)
        acc = acc + load_fr(m.coeff + k) * load_fr(x + m.col[k]);
    store_fr(y + r, acc);
}
// same, one warp per row (rows of the transposed matrices can hold millions of entries: the
// column of the constant ONE)
__global__ void spmv_warp_kernel(DevCsr m, const Fr* x, const Fr* extra, Fr* y) {
    uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= m.num_rows) return;
    Fr acc = Fr::zero();
    for (uint32_t k = m.row_ptr[r] + lane; k < m.row_ptr[r + 1]; k += 32)
        acc = acc + load_fr(m.coeff + k) * load_fr(x + m.col[k]);
    for (int o = 16; o > 0; o >>= 1) {
        Fr other;
        for (int j = 0; j < 8; j++) other.l[j] = __shfl_down_sync(0xffffffffu, acc.l[j], o);
        acc = acc + other;
    }
    if (lane == 0) {
        if (extra) acc = acc + load_fr(extra + r);
        store_fr(y + r, acc);
    }
}
// density maps (prover.rs:34-42): a variable's bit is set when it appears in a term, whatever the
// coefficient.  in_bits may be NULL (the A query: inputs have full density).
__global__ void density_kernel(const uint32_t* col, uint32_t nnz, uint32_t num_inputs, uint32_t* in_bits,
                               uint32_t* aux_bits) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    uint32_t c = col[k];
    if (c < num_inputs) {
        if (in_bits) atomicOr(in_bits + (c >> 5), 1u << (c & 31));
    } else {
        c -= num_inputs;
        atomicOr(aux_bits + (c >> 5), 1u << (c & 31));
    }
}
// out[i] = a[i] * k
__global__ void fr_scale_to_kernel(const Fr* a, const Fr* k, size_t n, int canon, Fr* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr v = load_fr(a + i) * *k;
    store_fr(out + i, canon ? v.from_mont() : v);
}
// ext[i] = (beta u[i] + alpha v[i] + w[i]) * inv   (generator.rs:502-510), canonical for the point mult
__global__ void ext_scalar_kernel(const Fr* u, const Fr* v, const Fr* w, const Fr* consts, size_t n, Fr* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr e = load_fr(u + i) * consts[1] + load_fr(v + i) * consts[0] + load_fr(w + i);   // consts: alpha, beta, inv
    store_fr(out + i, (e * consts[2]).from_mont());
}
// keygen constants: in[0..5) = alpha, beta, gamma, delta, tau (Montgomery); out: see enum below
enum { KG_ALPHA = 0, KG_BETA, KG_GAMMA_INV, KG_DELTA_INV, KG_TAU, KG_H_COEFF, KG_COUNT };
__global__ void keygen_consts_kernel(const Fr* in, uint32_t logm, Fr* out, uint32_t* bad) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Fr gamma = in[2], delta = in[3], tau = in[4];
    if (gamma.is_zero() || delta.is_zero()) *bad = 1;               // generator.rs:330-345
    Fr t = tau;
    for (uint32_t j = 0; j < logm; j++) t = t.sqr();
    Fr z = t - Fr::one();                                           // z(tau) = tau^m - 1
    Fr dinv = delta.inv();
    out[KG_ALPHA] = in[0];
    out[KG_BETA] = in[1];
    out[KG_GAMMA_INV] = gamma.inv();
    out[KG_DELTA_INV] = dinv;
    out[KG_TAU] = tau;
    out[KG_H_COEFF] = z * dinv;                                     // t(tau) / delta, generator.rs:368-369
}
// flags[i] = scalar i is non-zero; zero scalars give identity points that are filtered (generator.rs:618-632)
__global__ void nonzero_flags_kernel(const Fr* s, size_t n, uint32_t* flags) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = load_fr(s + i).is_zero() ? 0u : 1u;
}
template <int WORDS>
__global__ void compact_kernel(const uint32_t* src, const uint32_t* flags, const uint32_t* pos, size_t n, uint32_t* dst) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flags[i]) return;
    for (int j = 0; j < WORDS; j++) dst[(size_t)pos[i] * WORDS + j] = src[i * WORDS + j];
}
__global__ void any_zero_kernel(const Fr* s, size_t n, uint32_t* flag) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && load_fr(s + i).is_zero()) *flag = 1;
}

namespace {

struct UploadedCsr {
    uint32_t* row_ptr = nullptr;
    uint32_t* col = nullptr;
    Fr* coeff = nullptr;
    DevCsr view{};
    size_t nnz = 0;
    void release() {
        cudaFree(row_ptr); cudaFree(col); cudaFree(coeff);
        row_ptr = col = nullptr; coeff = nullptr;
    }
};

int upload_csr(bmpc_ctx* ctx, const bmpc_csr* m, cudaStream_t st, UploadedCsr* out) {
    out->nnz = m->nnz;
    CK(cudaMalloc(&out->row_ptr, (m->num_rows + 1) * 4));
    CK(cudaMalloc(&out->col, (m->nnz ? m->nnz : 1) * 4));
    CK(cudaMalloc(&out->coeff, (m->nnz ? m->nnz : 1) * sizeof(Fr)));
    CK(cudaMemcpyAsync(out->row_ptr, m->row_ptr, (m->num_rows + 1) * 4, cudaMemcpyHostToDevice, st));
    if (m->nnz) {
        CK(cudaMemcpyAsync(out->col, m->col, m->nnz * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(out->coeff, m->coeff, m->nnz * 32, cudaMemcpyHostToDevice, st));
    }
    out->view = DevCsr{out->row_ptr, out->col, out->coeff, (uint32_t)m->num_rows};
    return BMPC_OK;
}

int spmv(bmpc_ctx* ctx, const UploadedCsr& m, const Fr* x, const Fr* extra, Fr* y, bool warp_per_row, cudaStream_t st) {
    uint32_t rows = m.view.num_rows;
    if (!rows) return BMPC_OK;
    if (warp_per_row) LAUNCH(ctx, spmv_warp_kernel, (rows * 32 + 255) / 256, 256, 0, st, m.view, x, extra, y);
    else LAUNCH(ctx, spmv_thread_kernel, (rows + 127) / 128, 128, 0, st, m.view, x, extra, y);
    return BMPC_OK;
}

}  // namespace
}  // namespace bmpc

using namespace bmpc;

extern "C" {

int bmpc_r1cs_eval(bmpc_ctx* ctx, const bmpc_csr* A, const bmpc_csr* B, const bmpc_csr* C, size_t num_inputs,
                   size_t num_aux, const uint64_t* input_assignment, const uint64_t* aux_assignment,
                   uint64_t* a_out, uint64_t* b_out, uint64_t* c_out, uint64_t* a_aux_density,
                   uint64_t* b_input_density, uint64_t* b_aux_density) {
    if (!ctx || !A || !B || !C || !input_assignment || !a_out || !b_out || !c_out || !a_aux_density ||
        !b_input_density || !b_aux_density || (!aux_assignment && num_aux))
        return BMPC_ERR_INVALID;
    if (A->num_rows != B->num_rows || A->num_rows != C->num_rows) return BMPC_ERR_LENGTH_MISMATCH;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = ctx->own_stream;
    StreamScope ss(ctx, st);
    const size_t nc = A->num_rows, nv = num_inputs + num_aux, total = nc + num_inputs;
    const size_t wi = (num_inputs + 63) / 64, wa = (num_aux + 63) / 64;
    UploadedCsr m[3];
    const bmpc_csr* src[3] = {A, B, C};
    Fr *d_x = nullptr, *d_y = nullptr;
    uint32_t* d_bits = nullptr;
    auto cleanup = [&]() {
        cudaStreamSynchronize(st);
        for (auto& u : m) u.release();
        cudaFree(d_x); cudaFree(d_y); cudaFree(d_bits);
    };
    int rc = BMPC_OK;
    for (int k = 0; k < 3 && rc == BMPC_OK; k++) rc = upload_csr(ctx, src[k], st, &m[k]);
    if (rc) { cleanup(); return rc; }
#define RK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); cleanup(); return BMPC_ERR_CUDA; } } while (0)
    RK(cudaMalloc(&d_x, (nv ? nv : 1) * sizeof(Fr)));
    RK(cudaMalloc(&d_y, 3 * (total ? total : 1) * sizeof(Fr)));
    RK(cudaMalloc(&d_bits, (2 * wa + wi + 3) * 8));
    RK(cudaMemsetAsync(d_bits, 0, (2 * wa + wi + 3) * 8, st));
    RK(cudaMemcpyAsync(d_x, input_assignment, num_inputs * 32, cudaMemcpyHostToDevice, st));
    if (num_aux) RK(cudaMemcpyAsync(d_x + num_inputs, aux_assignment, num_aux * 32, cudaMemcpyHostToDevice, st));
    RK(cudaMemsetAsync(d_y, 0, 3 * (total ? total : 1) * sizeof(Fr), st));
    Fr* y[3] = {d_y, d_y + total, d_y + 2 * total};
    for (int k = 0; k < 3; k++) {
        rc = spmv(ctx, m[k], d_x, nullptr, y[k], false, st);
        if (rc) { cleanup(); return rc; }
    }
    // the `input_i * 0 = 0` rows appended by create_proof (prover.rs:202-204): a = input_i, b = c = 0
    RK(cudaMemcpyAsync(y[0] + nc, d_x, num_inputs * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
    uint32_t* bits_a_aux = d_bits;
    uint32_t* bits_b_in = d_bits + 2 * (wa + 1);
    uint32_t* bits_b_aux = bits_b_in + 2 * (wi + 1);
    if (A->nnz) LAUNCH(ctx, density_kernel, (uint32_t)((A->nnz + 255) / 256), 256, 0, st, m[0].col, (uint32_t)A->nnz, (uint32_t)num_inputs, (uint32_t*)nullptr, bits_a_aux);
    if (B->nnz) LAUNCH(ctx, density_kernel, (uint32_t)((B->nnz + 255) / 256), 256, 0, st, m[1].col, (uint32_t)B->nnz, (uint32_t)num_inputs, bits_b_in, bits_b_aux);
    uint64_t* outs[3] = {a_out, b_out, c_out};
    for (int k = 0; k < 3; k++) RK(cudaMemcpyAsync(outs[k], y[k], total * 32, cudaMemcpyDeviceToHost, st));
    if (wa) RK(cudaMemcpyAsync(a_aux_density, bits_a_aux, wa * 8, cudaMemcpyDeviceToHost, st));
    RK(cudaMemcpyAsync(b_input_density, bits_b_in, wi * 8, cudaMemcpyDeviceToHost, st));
    if (wa) RK(cudaMemcpyAsync(b_aux_density, bits_b_aux, wa * 8, cudaMemcpyDeviceToHost, st));
    RK(cudaStreamSynchronize(st));
#undef RK
    cleanup();
    return BMPC_OK;
}

int bmpc_generate_parameters(bmpc_ctx* ctx, const bmpc_csr* At, const bmpc_csr* Bt, const bmpc_csr* Ct,
                             size_t num_inputs, size_t num_aux, size_t num_constraints, const uint8_t g1[96],
                             const uint8_t g2[192], const uint64_t alpha[4], const uint64_t beta[4],
                             const uint64_t gamma[4], const uint64_t delta[4], const uint64_t tau[4],
                             bmpc_parameters* out) {
    if (!ctx || !At || !Bt || !Ct || !g1 || !g2 || !alpha || !beta || !gamma || !delta || !tau || !out)
        return BMPC_ERR_INVALID;
    const size_t nv = num_inputs + num_aux;
    if (At->num_rows != nv || Bt->num_rows != nv || Ct->num_rows != nv) return BMPC_ERR_LENGTH_MISMATCH;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard dg(ctx->device);
    cudaStream_t st = ctx->own_stream;
    StreamScope ss(ctx, st);
    memset(out, 0, sizeof(*out));
    // domain over the user constraints plus one `x * 0 = 0` row per input (generator.rs:273-275,294-297)
    const size_t nc = num_constraints + num_inputs;
    size_t m = 1;
    uint32_t exp = 0;
    while (m < nc) {
        m *= 2;
        exp++;
        if (exp >= 32) return BMPC_ERR_DEGREE_TOO_LARGE;
    }
    UploadedCsr mt[3];
    const bmpc_csr* src[3] = {At, Bt, Ct};
    Fr *d_in = nullptr, *d_k = nullptr, *d_pow = nullptr, *d_uvw = nullptr, *d_s = nullptr;
    uint32_t *d_flags = nullptr, *d_pos = nullptr, *d_chunks = nullptr, *d_bad = nullptr;
    void *d_base1 = nullptr, *d_base2 = nullptr, *d_tab1 = nullptr, *d_tab2 = nullptr, *d_pts = nullptr;
    uint8_t* d_raw = nullptr;
    bmpc_bases* made[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    auto cleanup = [&]() {
        cudaStreamSynchronize(st);
        for (auto& u : mt) u.release();
        cudaFree(d_in); cudaFree(d_k); cudaFree(d_pow); cudaFree(d_uvw); cudaFree(d_s);
        cudaFree(d_flags); cudaFree(d_pos); cudaFree(d_chunks); cudaFree(d_bad);
        cudaFree(d_base1); cudaFree(d_base2); cudaFree(d_tab1); cudaFree(d_tab2); cudaFree(d_pts); cudaFree(d_raw);
    };
    auto fail = [&](int code) {
        cleanup();
        for (auto b : made)
            if (b) bmpc_bases_free(ctx, b);
        memset(out, 0, sizeof(*out));
        return code;
    };
#define RK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); return fail(BMPC_ERR_CUDA); } } while (0)
#define RC(expr) do { int rc_ = (expr); if (rc_ != BMPC_OK) return fail(rc_); } while (0)
    for (int k = 0; k < 3; k++) RC(upload_csr(ctx, src[k], st, &mt[k]));
    RK(cudaMalloc(&d_in, 5 * sizeof(Fr)));
    RK(cudaMalloc(&d_k, KG_COUNT * sizeof(Fr)));
    RK(cudaMalloc(&d_bad, 16));
    RK(cudaMemsetAsync(d_bad, 0, 16, st));
    const uint64_t* sc[5] = {alpha, beta, gamma, delta, tau};
    for (int k = 0; k < 5; k++) RK(cudaMemcpyAsync(d_in + k, sc[k], 32, cudaMemcpyHostToDevice, st));
    LAUNCH(ctx, keygen_consts_kernel, 1, 1, 0, st, (const Fr*)d_in, exp, d_k, d_bad);
    uint32_t h_bad[4] = {0, 0, 0, 0};
    RK(cudaMemcpyAsync(h_bad, d_bad, 16, cudaMemcpyDeviceToHost, st));
    RK(cudaStreamSynchronize(st));
    if (h_bad[0]) return fail(BMPC_ERR_UNEXPECTED_IDENTITY);                         // gamma or delta not invertible

    // powers of tau (generator.rs:351-366), H query scalars tau^i t(tau)/delta (:372-397)
    const size_t big = m > nv ? m : nv;
    RK(cudaMalloc(&d_pow, m * sizeof(Fr)));
    RK(cudaMalloc(&d_s, (big ? big : 1) * sizeof(Fr)));
    RK(cudaMalloc(&d_uvw, 3 * (nv ? nv : 1) * sizeof(Fr)));
    LAUNCH(ctx, tau_powers_kernel, (uint32_t)((m + 127) / 128), 128, 0, st, (const Fr*)(d_k + KG_TAU), (uint32_t)m, d_pow);
    // fixed-base tables for g1, g2
    const size_t x1 = sizeof(G1XYZZ), x2 = sizeof(G2XYZZ);
    RK(cudaMalloc(&d_raw, 288));
    RK(cudaMalloc(&d_base1, 96)); RK(cudaMalloc(&d_base2, 192));
    RK(cudaMalloc(&d_tab1, 32 * 256 * x1)); RK(cudaMalloc(&d_tab2, 32 * 256 * x2));
    RK(cudaMemcpyAsync(d_raw, g1, 96, cudaMemcpyHostToDevice, st));
    RK(cudaMemcpyAsync(d_raw + 96, g2, 192, cudaMemcpyHostToDevice, st));
    RC(GroupOps<Fp>::decode(ctx, d_raw, 96, 1, d_base1, st));
    RC(GroupOps<Fp2>::decode(ctx, d_raw + 96, 192, 1, d_base2, st));
    RK(cudaMalloc(&d_pts, (big > 8 ? big : 8) * 192));      // also the 1152-byte vk encode scratch
    RK(cudaMalloc(&d_flags, (nv + 1) * 4)); RK(cudaMalloc(&d_pos, (nv + 2) * 4));
    RK(cudaMalloc(&d_chunks, (nv / 1024 + 4) * 4));

    // helper: scalars (canonical, device) -> points of `group` -> new resident bases (optionally compacted)
    auto make_bases = [&](int group, const Fr* scalars_canon, size_t n, const uint32_t* keep_pos, const uint32_t* keep_flags,
                          size_t n_keep, bmpc_bases** dst) -> int {
        const size_t pb = group == BMPC_G1 ? 96 : 192;
        int rc = group == BMPC_G1
                     ? GroupOps<Fp>::fixed_base_mul(ctx, d_base1, d_tab1, (const uint32_t*)scalars_canon, n, d_pts, st)
                     : GroupOps<Fp2>::fixed_base_mul(ctx, d_base2, d_tab2, (const uint32_t*)scalars_canon, n, d_pts, st);
        if (rc) return rc;
        bmpc_bases* b = new bmpc_bases();
        b->group = group;
        b->n = keep_pos ? n_keep : n;
        if (cudaMalloc(&b->d_points, (b->n ? b->n : 1) * pb) != cudaSuccess) { delete b; return BMPC_ERR_CUDA; }
        if (keep_pos) {
            if (n) {
                if (group == BMPC_G1) compact_kernel<24><<<(uint32_t)((n + 127) / 128), 128, 0, st>>>((const uint32_t*)d_pts, keep_flags, keep_pos, n, (uint32_t*)b->d_points);
                else compact_kernel<48><<<(uint32_t)((n + 127) / 128), 128, 0, st>>>((const uint32_t*)d_pts, keep_flags, keep_pos, n, (uint32_t*)b->d_points);
                ctx->launches++;
                cudaError_t ce = cudaGetLastError();
                if (ce != cudaSuccess) { ctx->err = std::string("compact_kernel: ") + cudaGetErrorString(ce); bmpc_bases_free(ctx, b); return BMPC_ERR_CUDA; }
            }
        } else if (n) {
            cudaMemcpyAsync(b->d_points, d_pts, n * pb, cudaMemcpyDeviceToDevice, st);
        }
        size_t nw = (b->n + 31) / 32 + 1;
        if (cudaMalloc(&b->d_inf, nw * 4) != cudaSuccess) { bmpc_bases_free(ctx, b); return BMPC_ERR_CUDA; }
        cudaMemsetAsync(b->d_inf, 0, nw * 4, st);
        rc = group == BMPC_G1 ? GroupOps<Fp>::inf_bitmap(ctx, b->d_points, b->n, b->d_inf, st)
                              : GroupOps<Fp2>::inf_bitmap(ctx, b->d_points, b->n, b->d_inf, st);
        if (rc) { bmpc_bases_free(ctx, b); return rc; }
        cudaStreamSynchronize(st);
        *dst = b;
        return BMPC_OK;
    };

    // H query: h[i] = g1 * (tau^i * t(tau)/delta), i < m - 1
    if (m > 1) LAUNCH(ctx, fr_scale_to_kernel, (uint32_t)((m - 1 + 255) / 256), 256, 0, st, (const Fr*)d_pow, (const Fr*)(d_k + KG_H_COEFF), m - 1, 1, d_s);
    RC(make_bases(BMPC_G1, d_s, m - 1, nullptr, nullptr, 0, &made[0]));

    // Lagrange coefficients: powers_of_tau.ifft() (generator.rs:401)
    RC(ntt_dev_locked(ctx, d_pow, exp, BMPC_IFFT, st));
    // u = At L, v = Bt L, w = Ct L per variable (eval_at_tau, :471-490).  The input rows
    // `x_i * 0 = 0` add L[num_constraints + i] to u_i (:273-275).
    Fr *d_u = d_uvw, *d_v = d_uvw + nv, *d_w = d_uvw + 2 * nv;
    RK(cudaMemsetAsync(d_s, 0, (nv ? nv : 1) * sizeof(Fr), st));
    RK(cudaMemcpyAsync(d_s, d_pow + num_constraints, num_inputs * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
    RC(spmv(ctx, mt[0], d_pow, d_s, d_u, true, st));
    RC(spmv(ctx, mt[1], d_pow, nullptr, d_v, true, st));
    RC(spmv(ctx, mt[2], d_pow, nullptr, d_w, true, st));

    // ext scalars: IC = (beta u + alpha v + w)/gamma for inputs, L = ... /delta for aux (:502-512)
    Fr* d_c3;
    RK(cudaMalloc(&d_c3, 3 * sizeof(Fr)));
    for (int part = 0; part < 2; part++) {
        size_t lo = part ? num_inputs : 0, cnt = part ? num_aux : num_inputs;
        RK(cudaMemcpyAsync(d_c3, d_k + KG_ALPHA, 2 * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
        RK(cudaMemcpyAsync(d_c3 + 2, d_k + (part ? KG_DELTA_INV : KG_GAMMA_INV), sizeof(Fr), cudaMemcpyDeviceToDevice, st));
        if (cnt) LAUNCH(ctx, ext_scalar_kernel, (uint32_t)((cnt + 255) / 256), 256, 0, st, (const Fr*)(d_u + lo), (const Fr*)(d_v + lo), (const Fr*)(d_w + lo), (const Fr*)d_c3, cnt, d_s);
        if (part) {   // unconstrained aux variable -> identity in L (:584-590)
            RK(cudaMemsetAsync(d_bad, 0, 16, st));
            if (cnt) LAUNCH(ctx, any_zero_kernel, (uint32_t)((cnt + 255) / 256), 256, 0, st, (const Fr*)d_s, cnt, d_bad);
            RK(cudaMemcpyAsync(h_bad, d_bad, 16, cudaMemcpyDeviceToHost, st));
            RK(cudaStreamSynchronize(st));
            if (h_bad[0]) { cudaFree(d_c3); ctx->err = "UnconstrainedVariable"; return fail(BMPC_ERR_INVALID_DATA); }
        }
        RC(make_bases(BMPC_G1, d_s, cnt, nullptr, nullptr, 0, &made[part ? 1 : 6]));   // l / ic
    }
    cudaFree(d_c3);

    // A query (G1, u) and B queries (G1 + G2, v): zero polynomials are omitted (:492-500,618-632)
    for (int q = 0; q < 2; q++) {
        const Fr* sv = q ? d_v : d_u;
        if (nv) LAUNCH(ctx, nonzero_flags_kernel, (uint32_t)((nv + 255) / 256), 256, 0, st, sv, nv, d_flags);
        // exclusive scan of the flags (reuse the MSM scan through a tiny single-block loop)
        std::vector<uint32_t> h_flags(nv + 1), h_pos(nv + 1);
        RK(cudaMemcpyAsync(h_flags.data(), d_flags, nv * 4, cudaMemcpyDeviceToHost, st));
        RK(cudaStreamSynchronize(st));
        uint32_t run = 0;
        for (size_t i = 0; i < nv; i++) { h_pos[i] = run; run += h_flags[i]; }   // index bookkeeping only
        RK(cudaMemcpyAsync(d_pos, h_pos.data(), (nv ? nv : 1) * 4, cudaMemcpyHostToDevice, st));
        RC(fr_pointwise(ctx, 2, const_cast<Fr*>(sv), nullptr, nv, st));            // to canonical, in place
        RC(make_bases(BMPC_G1, sv, nv, d_pos, d_flags, run, &made[q ? 3 : 2]));     // a / b_g1
        if (q) RC(make_bases(BMPC_G2, sv, nv, d_pos, d_flags, run, &made[4]));      // b_g2
    }

    // verifying key (:594-604)
    {
        Fr* d_vk = d_s;   // alpha, beta, gamma, delta canonical
        RK(cudaMemcpyAsync(d_vk, d_in, 4 * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
        RC(fr_pointwise(ctx, 2, d_vk, nullptr, 4, st));
        bmpc_bases *v1 = nullptr, *v2 = nullptr;
        RC(make_bases(BMPC_G1, d_vk, 4, nullptr, nullptr, 0, &v1));
        int rc2 = make_bases(BMPC_G2, d_vk, 4, nullptr, nullptr, 0, &v2);
        if (rc2) { bmpc_bases_free(ctx, v1); return fail(rc2); }
        uint8_t b1[4 * 96], b2[4 * 192];
        uint8_t* d_enc = reinterpret_cast<uint8_t*>(d_pts);       // free again: reuse as encode scratch
        int ra = GroupOps<Fp>::encode(ctx, v1->d_points, 4, d_enc, st);
        int rb = GroupOps<Fp2>::encode(ctx, v2->d_points, 4, d_enc + 384, st);
        if (!ra && !rb) {
            cudaMemcpyAsync(b1, d_enc, 384, cudaMemcpyDeviceToHost, st);
            cudaMemcpyAsync(b2, d_enc + 384, 768, cudaMemcpyDeviceToHost, st);
            cudaStreamSynchronize(st);
        }
        bmpc_bases_free(ctx, v1); bmpc_bases_free(ctx, v2);
        if (ra || rb) return fail(ra ? ra : rb);
        memcpy(out->p.alpha_g1, b1, 96); memcpy(out->p.beta_g1, b1 + 96, 96); memcpy(out->p.delta_g1, b1 + 288, 96);
        memcpy(out->p.beta_g2, b2 + 192, 192); memcpy(out->gamma_g2, b2 + 384, 192); memcpy(out->p.delta_g2, b2 + 576, 192);
    }
    out->p.h = made[0]; out->p.l = made[1]; out->p.a = made[2]; out->p.b_g1 = made[3]; out->p.b_g2 = made[4];
    out->ic = made[6];
    cleanup();
    return BMPC_OK;
#undef RK
#undef RC
}

}  // extern "C"
