// C++ host mirror (include/bellman_b200.hpp) exercised the way the reference's own tests
// exercise the Rust API: test_with_bls12 (multiexp.rs:283-327), fft_composition
// (domain.rs:427-463), the density-length assert (multiexp.rs:273-278) and Source EOF (:55-61).
// Expectations come from known discrete logs (sum k_i s_i) * G computed through the same library's
// fixed-base path on ONE point, so no CPU arithmetic is needed here.  Built and run by
// tests/test_gpu_cpp_mirror.py on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <random>

#include "../../include/bellman_b200.hpp"

using namespace bellman;

static const char* G1_GEN_HEX =
    "17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb"
    "08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1";

static std::vector<uint8_t> unhex(const char* h) {
    std::vector<uint8_t> out;
    for (size_t i = 0; h[i] && h[i + 1]; i += 2) {
        unsigned v;
        sscanf(h + i, "%2x", &v);
        out.push_back((uint8_t)v);
    }
    return out;
}
#define REQUIRE(c) do { if (!(c)) { fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main() {
    Worker worker(0);
    auto gen = unhex(G1_GEN_HEX);
    std::mt19937_64 rng(7);

    // ---- test_with_bls12: sum base_i * exp_i == multiexp(FullDensity)
    const size_t n = 1 << 12;
    std::vector<Scalar> k(n), s(n);
    unsigned __int128 dot = 0;
    for (size_t i = 0; i < n; i++) {
        uint64_t ki = rng() >> 34, si = rng() >> 34;      // < 2^30 each: the dot product fits in 128 bits
        k[i] = {ki, 0, 0, 0};
        s[i] = {si, 0, 0, 0};
        dot += (unsigned __int128)ki * si;
    }
    auto bases = Bases::fixed_base_mul(worker, BMPC_G1, gen.data(), k);
    std::vector<Scalar> dotv = {{(uint64_t)dot, (uint64_t)(dot >> 64), 0, 0}};
    auto expect = Bases::fixed_base_mul(worker, BMPC_G1, gen.data(), dotv)->read(0, 1);
    auto got = multiexp(worker, Source{bases, 0}, FullDensity{}, s).wait();
    REQUIRE(got == expect);
    bases->precompute();
    REQUIRE(multiexp(worker, Source{bases, 0}, FullDensity{}, s).wait() == expect);

    // ---- density map + offset: only even positions are dense, bases start at 3
    DensityTracker dens;
    unsigned __int128 dot2 = 0;
    size_t rank = 0;
    for (size_t i = 0; i < n - 3; i++) dens.add_element();
    std::vector<Scalar> s2(s.begin(), s.begin() + (n - 3));
    for (size_t i = 0; i < n - 3; i++)
        if (i % 2 == 0) { dens.inc(i); dot2 += (unsigned __int128)k[3 + rank][0] * s2[i][0]; rank++; }
    REQUIRE(dens.get_total_density() == rank);
    std::vector<Scalar> dot2v = {{(uint64_t)dot2, (uint64_t)(dot2 >> 64), 0, 0}};
    auto expect2 = Bases::fixed_base_mul(worker, BMPC_G1, gen.data(), dot2v)->read(0, 1);
    REQUIRE(multiexp(worker, Source{bases, 3}, dens, s2).wait() == expect2);

    // ---- the reference's assert on density length, and Source EOF
    bool threw = false;
    try { multiexp(worker, Source{bases, 0}, dens, s); } catch (const std::logic_error&) { threw = true; }
    REQUIRE(threw);
    threw = false;
    try { multiexp(worker, Source{bases, 5}, FullDensity{}, s).wait(); } catch (const UnexpectedEof&) { threw = true; }
    REQUIRE(threw);

    // ---- list_mul_matrix (mpc.rs:416-457) and per-element scalar multiplication (:647-706):
    // row i = sum_j m_ij * (k_idx G) == (sum_j m_ij k_idx) G; rows after the first empty one stay O
    {
        const size_t ln = 24;
        std::vector<Scalar> lk(k.begin(), k.begin() + ln);
        auto l1 = Bases::fixed_base_mul(worker, BMPC_G1, gen.data(), lk);
        SparseMatrix m(10);
        std::vector<Scalar> rowdot(ln, Scalar{0, 0, 0, 0});
        for (size_t i = 0; i < m.size(); i++) {
            if (i == 7) continue;                                   // the reference stops here
            unsigned __int128 acc = 0;
            for (size_t j = 0; j < 1 + i % 4; j++) {
                uint64_t cf = rng() >> 34;
                size_t idx = rng() % ln;
                m[i].push_back({Scalar{cf, 0, 0, 0}, idx});
                acc += (unsigned __int128)cf * lk[idx][0];
            }
            if (i < 7) rowdot[i] = {(uint64_t)acc, (uint64_t)(acc >> 64), 0, 0};
        }
        auto res = list_mul_matrix(worker, *l1, *l1, m);           // the same G1 list in both slots
        auto want = Bases::fixed_base_mul(worker, BMPC_G1, gen.data(), rowdot)->read(0, ln);
        REQUIRE(res.first->len() == ln && res.first->read(0, ln) == want);
        REQUIRE(res.second->read(0, ln) == want);
        bool oob = false;
        try { list_mul_matrix(worker, *l1, *l1, SparseMatrix{{{Scalar{1, 0, 0, 0}, ln}}}); } catch (const std::logic_error&) { oob = true; }
        REQUIRE(oob);
        std::vector<Scalar> mult(ln), prod(ln);
        for (size_t i = 0; i < ln; i++) {
            uint64_t mi = rng() >> 34;
            unsigned __int128 pr = (unsigned __int128)mi * lk[i][0];
            mult[i] = {mi, 0, 0, 0};
            prod[i] = {(uint64_t)pr, (uint64_t)(pr >> 64), 0, 0};
        }
        REQUIRE(l1->scalar_mul(mult)->read(0, ln) == Bases::fixed_base_mul(worker, BMPC_G1, gen.data(), prod)->read(0, ln));
    }

    // ---- several multiexps in flight before the first wait() (prover.rs:233-307 / multicore.rs:93-118)
    {
        auto w1 = multiexp(worker, Source{bases, 0}, FullDensity{}, s);
        auto w2 = multiexp(worker, Source{bases, 3}, dens, s2);
        auto w3 = multiexp(worker, Source{bases, 0}, FullDensity{}, s);
        REQUIRE(w3.wait() == expect);
        REQUIRE(w2.wait() == expect2);
        REQUIRE(w1.wait() == expect);
    }

    // ---- all devices of the node behind ONE call (SURVEY 8b bmpc_ctx_create(devices, n)): here the
    // same GPU twice / three times, which runs the whole plan (base split, position cuts through the
    // density map, peer gather, fold) on a one-GPU box.  Same bytes as the single-device calls.
    for (int nd = 2; nd <= 3; nd++) {
        int devs[3] = {0, 0, 0};
        bmpc_multi* mw = nullptr;
        REQUIRE(bmpc_multi_create(devs, nd, &mw) == BMPC_OK && bmpc_multi_size(mw) == nd);
        auto raw = bases->read(0, n);
        bmpc_multi_bases* mb = nullptr;
        REQUIRE(bmpc_multi_bases_register(mw, BMPC_G1, raw.data(), n, 96, BMPC_FORM_UNCOMPRESSED_BE, &mb) == BMPC_OK);
        REQUIRE(bmpc_multi_bases_len(mb) == n);
        std::vector<uint8_t> out(96);
        REQUIRE(bmpc_multi_multiexp(mw, mb, 0, s[0].data(), n, nullptr, 0, out.data()) == BMPC_OK);
        REQUIRE(out == expect);
        REQUIRE(bmpc_multi_multiexp(mw, mb, 3, s2[0].data(), n - 3, dens.words(), n - 3, out.data()) == BMPC_OK);
        REQUIRE(out == expect2);
        REQUIRE(bmpc_multi_bases_precompute(mw, mb, 0) == BMPC_OK);
        REQUIRE(bmpc_multi_multiexp(mw, mb, 0, s[0].data(), n, nullptr, 0, out.data()) == BMPC_OK);
        REQUIRE(out == expect);
        REQUIRE(bmpc_multi_multiexp(mw, mb, 5, s[0].data(), n, nullptr, 0, out.data()) == BMPC_ERR_UNEXPECTED_EOF);
        REQUIRE(bmpc_multi_multiexp(mw, mb, 0, s[0].data(), n, dens.words(), n - 3, out.data()) == BMPC_ERR_LENGTH_MISMATCH);
        bmpc_multi_bases_free(mw, mb);
        bmpc_multi_destroy(mw);
    }

    // ---- fft_composition (domain.rs:427-463)
    for (unsigned logn = 0; logn < 11; logn++) {
        std::vector<Scalar> c(size_t(1) << logn);
        for (auto& v : c) v = {rng(), rng(), rng(), rng() >> 2};      // any value < 2^254 < q
        auto d = EvaluationDomain::from_coeffs(worker, c);
        d.ifft(worker); d.fft(worker);
        REQUIRE(d.into_coeffs() == c);
        d.coset_fft(worker); d.icoset_fft(worker);
        REQUIRE(d.into_coeffs() == c);
    }
    printf("cpp mirror ok\n");
    return 0;
}
