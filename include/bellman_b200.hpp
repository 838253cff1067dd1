// bellman_b200.hpp -- C++ host-side mirror of the reference's hot-path interface, header-only,
// over the C ABI in bellman_b200.h.  Same names, argument meaning and error behaviour as the Rust
// reference (doubiliu/bellman-mpc, /root/reference/bellman/src):
//
//   Worker / Waiter<T> ............................ multicore.rs:21-118
//   FullDensity / DensityTracker .................. multiexp.rs:88-157
//   multiexp(pool, bases, density_map, exponents) . multiexp.rs:254-281
//   EvaluationDomain .............................. domain.rs:21-189
//   Parameters / ProvingAssignment / create_proof . groth16/mod.rs:224-247, prover.rs:55-69,176-350
//   SynthesisError ................................ lib.rs:355-370
//
// This is what a C++ host (or the Rust shim of INTEGRATION.md) sits on; there is no CPU fallback.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "bellman_b200.h"

namespace bellman {

// ---- SynthesisError (lib.rs:355-370) ------------------------------------------------------
struct SynthesisError : std::runtime_error { using std::runtime_error::runtime_error; };
struct UnexpectedIdentity : SynthesisError { UnexpectedIdentity() : SynthesisError("UnexpectedIdentity") {} };
struct IoError : SynthesisError { using SynthesisError::SynthesisError; };
struct UnexpectedEof : IoError { UnexpectedEof() : IoError("expected more bases from source") {} };
struct PolynomialDegreeTooLarge : SynthesisError { PolynomialDegreeTooLarge() : SynthesisError("PolynomialDegreeTooLarge") {} };

inline void check(int st, const bmpc_ctx* ctx = nullptr) {
    switch (st) {
        case BMPC_OK: return;
        case BMPC_ERR_UNEXPECTED_IDENTITY: throw UnexpectedIdentity();
        case BMPC_ERR_UNEXPECTED_EOF: throw UnexpectedEof();
        case BMPC_ERR_DEGREE_TOO_LARGE: throw PolynomialDegreeTooLarge();
        case BMPC_ERR_LENGTH_MISMATCH: throw std::logic_error("length mismatch (the reference asserts)");
        case BMPC_ERR_CUDA: throw IoError(std::string("CUDA error: ") + (ctx ? bmpc_last_error(ctx) : ""));
        default: throw std::invalid_argument("bmpc: invalid argument");
    }
}

using Scalar = std::array<uint64_t, 4>;   // Fr: canonical (exponents) or Montgomery (coefficients)

// ---- Worker / Waiter (multicore.rs) ---------------------------------------------------------
class Worker {
public:
    explicit Worker(int device = 0) { check(bmpc_ctx_create(device, &ctx_)); }
    ~Worker() { bmpc_ctx_destroy(ctx_); }
    Worker(const Worker&) = delete;
    Worker& operator=(const Worker&) = delete;
    bmpc_ctx* ctx() const { return ctx_; }
private:
    bmpc_ctx* ctx_ = nullptr;
};

template <class T>
class Waiter {            // multicore.rs:93-118; wait() yields the Result (throws the error)
public:
    static Waiter done(T v) { Waiter w; w.value_ = std::move(v); return w; }
    static Waiter failed(std::exception_ptr e) { Waiter w; w.err_ = e; return w; }
    // work still in flight on the GPU: `finish` blocks for it (bmpc_waiter_wait) and yields the value
    static Waiter pending(std::function<T()> finish) { Waiter w; w.finish_ = std::move(finish); return w; }
    Waiter() = default;
    Waiter(Waiter&&) = default;
    Waiter& operator=(Waiter&&) = default;
    ~Waiter() { if (finish_) { try { finish_(); } catch (...) {} } }      // a dropped waiter gives its lane back
    T wait() {
        if (finish_) { auto f = std::move(finish_); finish_ = nullptr; return f(); }
        if (err_) std::rethrow_exception(err_);
        return std::move(value_);
    }
private:
    T value_{};
    std::exception_ptr err_;
    std::function<T()> finish_;
};

// ---- bases: (Arc<Vec<G::Affine>>, usize) SourceBuilder (multiexp.rs:45-86) --------------------
class Bases {
public:
    Bases(const Worker& w, int group, const uint8_t* uncompressed, size_t n) : w_(&w) {
        check(bmpc_bases_register(w.ctx(), group, uncompressed, n, 0, BMPC_FORM_UNCOMPRESSED_BE, &h_), w.ctx());
    }
    Bases(const Worker& w, bmpc_bases* h) : w_(&w), h_(h) {}
    ~Bases() { bmpc_bases_free(w_->ctx(), h_); }
    Bases(const Bases&) = delete;
    Bases& operator=(const Bases&) = delete;
    static std::shared_ptr<Bases> fixed_base_mul(const Worker& w, int group, const uint8_t* base,
                                                 const std::vector<Scalar>& k) {
        bmpc_bases* h = nullptr;
        check(bmpc_fixed_base_mul(w.ctx(), group, base, k.empty() ? nullptr : k[0].data(), k.size(), 0, &h), w.ctx());
        return std::make_shared<Bases>(w, h);
    }
    // mpc.rs:647-706 make_new_paramter / make_new_tau_paramter: element i times k[i], or every
    // element times k[0] (per_element = false)
    std::shared_ptr<Bases> scalar_mul(const std::vector<Scalar>& k, bool per_element = true) const {
        bmpc_bases* h = nullptr;
        if (k.empty() || (per_element && k.size() != len())) throw std::logic_error("length mismatch");
        check(bmpc_batch_scalar_mul(w_->ctx(), h_, k[0].data(), per_element ? 1 : 0, &h), w_->ctx());
        return std::make_shared<Bases>(*w_, h);
    }
    void precompute(int window_bits = 0) { check(bmpc_bases_precompute(w_->ctx(), h_, window_bits), w_->ctx()); }
    size_t len() const { return bmpc_bases_len(h_); }
    int group() const { return bmpc_bases_group(h_); }
    std::vector<uint8_t> read(size_t start, size_t count) const {
        std::vector<uint8_t> out(count * (group() == BMPC_G1 ? 96 : 192));
        check(bmpc_bases_read(w_->ctx(), h_, start, count, out.data()), w_->ctx());
        return out;
    }
    const bmpc_bases* handle() const { return h_; }
private:
    const Worker* w_;
    bmpc_bases* h_ = nullptr;
};
using Source = std::pair<std::shared_ptr<Bases>, size_t>;

// ---- list_mul_matrix (groth16/mpc.rs:416-457) -------------------------------------------------
// result[i] = sum_j list[matrix[i][j].second] * matrix[i][j].first for the rows before the first
// empty one (the reference `break`s there), identity elsewhere; both results have the lists' length.
// Coefficients canonical.  An index out of range panics in the reference: std::logic_error here.
using SparseMatrix = std::vector<std::vector<std::pair<Scalar, size_t>>>;
inline std::pair<std::shared_ptr<Bases>, std::shared_ptr<Bases>>
list_mul_matrix(const Worker& w, const Bases& list_g1, const Bases& list_g2, const SparseMatrix& matrix) {
    std::vector<uint64_t> row_ptr(matrix.size() + 1, 0), coeffs;
    std::vector<uint32_t> cols;
    for (size_t i = 0; i < matrix.size(); i++) {
        row_ptr[i + 1] = row_ptr[i] + matrix[i].size();
        for (const auto& e : matrix[i]) {
            if (e.second > 0xffffffffull) throw std::logic_error("index out of bounds");
            cols.push_back((uint32_t)e.second);
            coeffs.insert(coeffs.end(), e.first.begin(), e.first.end());
        }
    }
    bmpc_bases *h1 = nullptr, *h2 = nullptr;
    check(bmpc_list_mul_matrix(w.ctx(), list_g1.handle(), row_ptr.data(), cols.data(), coeffs.data(), matrix.size(), &h1), w.ctx());
    auto r1 = std::make_shared<Bases>(w, h1);
    check(bmpc_list_mul_matrix(w.ctx(), list_g2.handle(), row_ptr.data(), cols.data(), coeffs.data(), matrix.size(), &h2), w.ctx());
    return {r1, std::make_shared<Bases>(w, h2)};
}

// ---- QueryDensity (multiexp.rs:88-157) --------------------------------------------------------
struct FullDensity {
    bool has_query_size() const { return false; }
    size_t get_query_size() const { return 0; }
    const uint64_t* words() const { return nullptr; }
};
class DensityTracker {
public:
    void add_element() { if (n_ % 64 == 0) bv_.push_back(0); n_++; }
    void inc(size_t idx) { if (idx >= n_) throw std::out_of_range("DensityTracker::inc"); bv_[idx / 64] |= uint64_t(1) << (idx % 64); }
    size_t get_total_density() const { size_t c = 0; for (uint64_t w : bv_) c += __builtin_popcountll(w); return c; }
    bool has_query_size() const { return true; }
    size_t get_query_size() const { return n_; }
    const uint64_t* words() const { static const uint64_t zero = 0; return bv_.empty() ? &zero : bv_.data(); }
private:
    std::vector<uint64_t> bv_;   // BitVec<Lsb0, usize> raw storage
    size_t n_ = 0;
};

// ---- multiexp (multiexp.rs:254-281) ------------------------------------------------------------
// exponents: canonical little-endian scalars.  The Waiter yields the uncompressed affine sum
// (G::to_affine().to_uncompressed(): 96 B for G1, 192 B for G2).
template <class D>
Waiter<std::vector<uint8_t>> multiexp(const Worker& pool, const Source& bases, const D& density_map,
                                      const std::vector<Scalar>& exponents) {
    if (density_map.has_query_size() && density_map.get_query_size() != exponents.size())
        throw std::logic_error("assertion failed: query_size == exponents.len()");      // multiexp.rs:273-278
    // Returns at once like the reference (the multiexp runs on a lane of the context); `exponents`
    // and the density map must outlive wait(), as the Arc<Vec<..>> arguments of the reference do.
    const size_t nbytes = bases.first->group() == BMPC_G1 ? 96 : 192;
    bmpc_waiter* h = nullptr;
    int st = bmpc_multiexp_async(pool.ctx(), bases.first->handle(), bases.second,
                                 exponents.empty() ? nullptr : exponents[0].data(), exponents.size(),
                                 density_map.words(), density_map.words() ? exponents.size() : 0, &h);
    try { check(st, pool.ctx()); } catch (...) { return Waiter<std::vector<uint8_t>>::failed(std::current_exception()); }
    bmpc_ctx* ctx = pool.ctx();
    auto keep = bases.first;
    return Waiter<std::vector<uint8_t>>::pending([h, ctx, nbytes, keep]() {
        std::vector<uint8_t> out(nbytes);
        check(bmpc_waiter_wait(h, out.data()), ctx);
        return out;
    });
}

// ---- EvaluationDomain<Fr, Scalar<Fr>> (domain.rs:21-189), coefficients resident in HBM ----------
class EvaluationDomain {
public:
    static EvaluationDomain from_coeffs(const Worker& w, const std::vector<Scalar>& coeffs) {
        bmpc_domain* d = nullptr;
        check(bmpc_domain_from_coeffs(w.ctx(), coeffs.empty() ? nullptr : coeffs[0].data(), coeffs.size(), &d), w.ctx());
        return EvaluationDomain(w, d);
    }
    EvaluationDomain(EvaluationDomain&& o) noexcept : w_(o.w_), d_(o.d_) { o.d_ = nullptr; }
    ~EvaluationDomain() { if (d_) bmpc_domain_free(w_->ctx(), d_); }
    size_t len() const { return bmpc_domain_len(d_); }
    std::vector<Scalar> into_coeffs() const {
        std::vector<Scalar> out(len());
        check(bmpc_domain_into_coeffs(w_->ctx(), d_, out[0].data()), w_->ctx());
        return out;
    }
    void fft(const Worker&) { t(BMPC_FFT); }
    void ifft(const Worker&) { t(BMPC_IFFT); }
    void coset_fft(const Worker&) { t(BMPC_COSET_FFT); }
    void icoset_fft(const Worker&) { t(BMPC_ICOSET_FFT); }
    void distribute_powers(const Worker&, const Scalar& g) { check(bmpc_domain_distribute_powers(w_->ctx(), d_, g.data(), nullptr), w_->ctx()); }
    Scalar z(const Scalar& tau) const { Scalar o; check(bmpc_domain_z(w_->ctx(), d_, tau.data(), o.data()), w_->ctx()); return o; }
    void divide_by_z_on_coset(const Worker&) { check(bmpc_domain_divide_by_z_on_coset(w_->ctx(), d_, nullptr), w_->ctx()); }
    void mul_assign(const Worker&, const EvaluationDomain& o) { check(bmpc_domain_mul_assign(w_->ctx(), d_, o.d_, nullptr), w_->ctx()); }
    void sub_assign(const Worker&, const EvaluationDomain& o) { check(bmpc_domain_sub_assign(w_->ctx(), d_, o.d_, nullptr), w_->ctx()); }
private:
    EvaluationDomain(const Worker& w, bmpc_domain* d) : w_(&w), d_(d) {}
    void t(int op) { check(bmpc_domain_transform(w_->ctx(), d_, op, nullptr), w_->ctx()); }
    const Worker* w_;
    bmpc_domain* d_;
};

// ---- groth16 (groth16/mod.rs:224-247, prover.rs) ------------------------------------------------
struct Parameters {
    std::shared_ptr<Bases> h, l, a, b_g1, b_g2;
    std::array<uint8_t, 96> alpha_g1{}, beta_g1{}, delta_g1{};
    std::array<uint8_t, 192> beta_g2{}, delta_g2{};
};
struct ProvingAssignment {           // prover.rs:55-69 after synthesis (Montgomery limbs)
    std::vector<Scalar> a, b, c, input_assignment, aux_assignment;
    DensityTracker a_aux_density, b_input_density, b_aux_density;
};
using Proof = std::array<uint8_t, 192>;

inline Proof create_proof(const Worker& w, const ProvingAssignment& s, const Parameters& p, const Scalar& r,
                          const Scalar& sv) {
    bmpc_params P;
    std::memset(&P, 0, sizeof(P));
    P.h = p.h->handle(); P.l = p.l->handle(); P.a = p.a->handle(); P.b_g1 = p.b_g1->handle(); P.b_g2 = p.b_g2->handle();
    std::memcpy(P.alpha_g1, p.alpha_g1.data(), 96); std::memcpy(P.beta_g1, p.beta_g1.data(), 96);
    std::memcpy(P.beta_g2, p.beta_g2.data(), 192); std::memcpy(P.delta_g1, p.delta_g1.data(), 96);
    std::memcpy(P.delta_g2, p.delta_g2.data(), 192);
    auto ptr = [](const std::vector<Scalar>& v) { return v.empty() ? nullptr : v[0].data(); };
    bmpc_assignment A{ptr(s.a), ptr(s.b), ptr(s.c), s.a.size(), ptr(s.input_assignment), s.input_assignment.size(),
                      ptr(s.aux_assignment), s.aux_assignment.size(), s.a_aux_density.words(),
                      s.b_input_density.words(), s.b_aux_density.words()};
    Proof out;
    check(bmpc_create_proof(w.ctx(), &P, &A, r.data(), sv.data(), out.data()), w.ctx());
    return out;
}

}  // namespace bellman
