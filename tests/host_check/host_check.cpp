// Host-side build of the device arithmetic headers (carry flag emulated) so the
// algorithms can be checked bit-for-bit against the big-integer oracle without a GPU.
// Test infrastructure only.
#include "../../bellman_mpc_b200/csrc/curve.cuh"
#include "../../bellman_mpc_b200/csrc/msm_affine.cuh"
#include "../../bellman_mpc_b200/csrc/msm_pairs.cuh"
#include <vector>
#include <string.h>
using namespace bmpc;

// One batched-affine accumulate job (msm_affine.cuh) on the host next to the XYZZ chain it
// replaces.  tasks: BMPC_AFF_G x {first sorted entry, length}; outputs: per task an affine point
// (tree result, after XYZZ -> affine) and the chain result.
template <class F>
static void affine_job_host(const uint32_t* bases_w, const uint32_t* sorted, const uint32_t* tasks, uint32_t L,
                            uint32_t* out_tree, uint32_t* out_chain) {
    const Affine<F>* bases = reinterpret_cast<const Affine<F>*>(bases_w);
    const uint32_t G = BMPC_AFF_G, HA = (L + 1) / 2, HB = (HA + 1) / 2;
    std::vector<Affine<F>> scratch((size_t)G * (HA + HB));
    std::vector<XYZZ<F>> partials(G);
    std::vector<F> pre(BMPC_AFF_K);
    AffJob<F> J;
    J.bases = bases; J.sorted = sorted;
    J.bufA = scratch.data(); J.bufB = scratch.data() + (size_t)G * HA; J.HA = HA; J.HB = HB; J.G = G;
    for (uint32_t g = 0; g < G; g++) { J.start[g] = tasks[2 * g]; J.len[g] = tasks[2 * g + 1]; J.slot[g] = g; }
    for (uint32_t g = 0; g < G; g++) partials[g] = XYZZ<F>::identity();
    aff_run_job<F>(J, pre.data(), (uint32_t)BMPC_AFF_K, partials.data(), SoloCoop());
    for (uint32_t g = 0; g < G; g++) {
        Affine<F> t = partials[g].to_affine();
        memcpy(out_tree + g * sizeof(Affine<F>) / 4, &t, sizeof(Affine<F>));
        XYZZ<F> acc = XYZZ<F>::identity();
        for (uint32_t j = 0; j < tasks[2 * g + 1]; j++) {
            uint32_t e = sorted[tasks[2 * g] + j];
            Affine<F> p = bases[e & 0x7fffffffu];
            if (e >> 31) p.y = p.y.neg();
            acc.add_affine(p);
        }
        Affine<F> c = acc.to_affine();
        memcpy(out_chain + g * sizeof(Affine<F>) / 4, &c, sizeof(Affine<F>));
    }
}

// The round-based pair accumulation (msm_pairs.cuh) on the host: the same list rule
// (pair_build_task) and the same forward / backward steps the kernel runs, chunks of `chunk` pairs
// sharing one inversion.  sorted: the padded entry array (every task an even number of entries,
// pad = BMPC_PAIR_PAD); tasks: ntasks x {first entry, padded length}.  out_pairs / out_chain: per task
// the affine sum by rounds and by the XYZZ chain.  Returns the number of additions listed.
template <class F>
static uint64_t pairs_host(const uint32_t* bases_w, const uint32_t* sorted, uint32_t nsorted, const uint32_t* tasks,
                           uint32_t ntasks, uint32_t R, uint32_t chunk, uint32_t* out_pairs, uint32_t* out_chain) {
    const Affine<F>* bases = reinterpret_cast<const Affine<F>*>(bases_w);
    // per-round scans over the tasks
    std::vector<std::vector<uint32_t>> pairoff(R, std::vector<uint32_t>(ntasks, 0));
    std::vector<uint32_t> totals(R + 1, 0), out_base(R + 1, 0);
    totals[0] = nsorted / 2;
    for (uint32_t r = 1; r < R; r++) {
        uint32_t acc = 0;
        for (uint32_t t = 0; t < ntasks; t++) { pairoff[r][t] = acc; acc += pair_count(tasks[2 * t + 1], r); }
        totals[r] = acc;
    }
    for (uint32_t r = 0; r < R; r++) out_base[r + 1] = out_base[r] + totals[r] + 3;   // slack like the host bounds
    std::vector<std::vector<PairIdx>> lists(R);
    for (uint32_t r = 1; r < R; r++) lists[r].resize(totals[r] + 1);
    std::vector<uint32_t> fin(ntasks);
    for (uint32_t t = 0; t < ntasks; t++) {
        uint32_t po[BMPC_PAIR_MAX_ROUNDS + 1];
        PairIdx* lp[BMPC_PAIR_MAX_ROUNDS + 1];
        for (uint32_t r = 0; r < R; r++) { po[r] = r ? pairoff[r][t] : 0; lp[r] = r ? lists[r].data() : nullptr; }
        fin[t] = pair_build_task(tasks[2 * t], tasks[2 * t + 1], R, po, out_base.data(), lp);
    }
    std::vector<Affine<F>> pool(out_base[R] + 1, Affine<F>::identity());
    uint64_t adds = 0;
    for (uint32_t r = 0; r < R; r++) {
        const uint32_t P = totals[r];
        for (uint32_t c0 = 0; c0 < P; c0 += chunk) {
            const uint32_t kt = P - c0 < chunk ? P - c0 : chunk;
            std::vector<F> pre(kt);
            std::vector<Affine<F>> Ps(kt), Qs(kt);
            F acc = F::one();
            for (uint32_t j = 0; j < kt; j++) {
                const uint32_t k = c0 + j;
                bool single = false;
                if (r == 0) {
                    uint32_t ea = sorted[2 * k], eb = sorted[2 * k + 1];
                    Ps[j] = bases[ea & 0x7fffffffu];
                    if (ea >> 31) Ps[j].y = Ps[j].y.neg();
                    single = eb == BMPC_PAIR_PAD;
                    if (single) Qs[j] = Affine<F>::identity();
                    else { Qs[j] = bases[eb & 0x7fffffffu]; if (eb >> 31) Qs[j].y = Qs[j].y.neg(); }
                } else {
                    Ps[j] = pool[lists[r][k].a];
                    Qs[j] = pool[lists[r][k].b];
                }
                F d;
                int q = pair_fwd_quick<F>(Ps[j].x, single ? Ps[j].x : Qs[j].x, single, d);
                bool use = q == 0;
                if (q == 2) use = aff_classify<F>(Ps[j], Qs[j], d) != 2;
                pre[j] = acc;
                if (use) { acc = acc * d; adds++; }
            }
            F inv = acc.inv();
            for (uint32_t j = kt; j-- > 0;) pool[out_base[r] + c0 + j] = pair_bwd_add<F>(Ps[j], Qs[j], pre[j], inv);
        }
    }
    for (uint32_t t = 0; t < ntasks; t++) {
        Affine<F> v = fin[t] == BMPC_PAIR_NONE ? Affine<F>::identity() : pool[fin[t]];
        memcpy(out_pairs + (size_t)t * sizeof(Affine<F>) / 4, &v, sizeof(Affine<F>));
        XYZZ<F> a = XYZZ<F>::identity();
        for (uint32_t j = 0; j < tasks[2 * t + 1]; j++) {
            uint32_t e = sorted[tasks[2 * t] + j];
            if (e == BMPC_PAIR_PAD) continue;
            Affine<F> p = bases[e & 0x7fffffffu];
            if (e >> 31) p.y = p.y.neg();
            a.add_affine(p);
        }
        Affine<F> cch = a.to_affine();
        memcpy(out_chain + (size_t)t * sizeof(Affine<F>) / 4, &cch, sizeof(Affine<F>));
    }
    return adds;
}

extern "C" {
uint64_t hc_pairs_g1(const uint32_t* bases, const uint32_t* sorted, uint32_t nsorted, const uint32_t* tasks,
                     uint32_t ntasks, uint32_t R, uint32_t chunk, uint32_t* out_pairs, uint32_t* out_chain) {
    return pairs_host<Fp>(bases, sorted, nsorted, tasks, ntasks, R, chunk, out_pairs, out_chain);
}
uint64_t hc_pairs_g2(const uint32_t* bases, const uint32_t* sorted, uint32_t nsorted, const uint32_t* tasks,
                     uint32_t ntasks, uint32_t R, uint32_t chunk, uint32_t* out_pairs, uint32_t* out_chain) {
    return pairs_host<Fp2>(bases, sorted, nsorted, tasks, ntasks, R, chunk, out_pairs, out_chain);
}
uint32_t hc_affine_g() { return BMPC_AFF_G; }
void hc_affine_job_g1(const uint32_t* bases, const uint32_t* sorted, const uint32_t* tasks, uint32_t L,
                      uint32_t* out_tree, uint32_t* out_chain) {
    affine_job_host<Fp>(bases, sorted, tasks, L, out_tree, out_chain);
}
void hc_affine_job_g2(const uint32_t* bases, const uint32_t* sorted, const uint32_t* tasks, uint32_t L,
                      uint32_t* out_tree, uint32_t* out_chain) {
    affine_job_host<Fp2>(bases, sorted, tasks, L, out_tree, out_chain);
}
// op: 0 mul, 1 add, 2 sub, 3 neg, 4 inv, 5 to_mont, 6 from_mont, 7 sqr, 8 inv_fermat
void hc_fr_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* r) {
    Fr x, y, z; memcpy(x.l, a, 32); memcpy(y.l, b, 32);
    switch (op) { case 0: z = x * y; break; case 1: z = x + y; break; case 2: z = x - y; break;
        case 3: z = x.neg(); break; case 4: z = x.inv(); break; case 5: z = x.to_mont(); break;
        case 6: z = x.from_mont(); break; case 8: z = x.inv_fermat(); break; default: z = x.sqr_redc(); }
    memcpy(r, z.l, 32);
}
void hc_fp_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* r) {
    Fp x, y, z; memcpy(x.l, a, 48); memcpy(y.l, b, 48);
    switch (op) { case 0: z = x * y; break; case 1: z = x + y; break; case 2: z = x - y; break;
        case 3: z = x.neg(); break; case 4: z = x.inv(); break; case 5: z = x.to_mont(); break;
        case 6: z = x.from_mont(); break; case 8: z = x.inv_fermat(); break; default: z = x.sqr_redc(); }
    memcpy(r, z.l, 48);
}
void hc_fp2_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* r) {
    Fp2 x, y, z; memcpy(&x, a, 96); memcpy(&y, b, 96);
    switch (op) { case 0: z = x * y; break; case 1: z = x + y; break; case 2: z = x - y; break;
        case 3: z = x.neg(); break; case 4: z = x.inv(); break; default: z = x.sqr(); }
    memcpy(r, &z, 96);
}
// G1 ops on affine Montgomery points (24 words each, (0,0) = identity)
// op: 0 add (via XYZZ add_affine), 1 full add (XYZZ+XYZZ with non-trivial Z), 2 double, 3 scalar mul by k[8]
void hc_g1_op(int op, const uint32_t* a, const uint32_t* b, const uint32_t* k, uint32_t* r) {
    G1Affine p, q; memcpy(&p, a, 96); memcpy(&q, b, 96);
    G1XYZZ acc = G1XYZZ::from_affine(p);
    if (op == 0) acc.add_affine(q);
    else if (op == 1) {
        // give both operands non-trivial denominators: P = (2P - P'), Q likewise
        G1XYZZ P2 = acc.dbl(); P2.add_affine(p.neg());
        G1XYZZ Q = G1XYZZ::from_affine(q); G1XYZZ Q2 = Q.dbl(); Q2.add_affine(q.neg());
        P2.add(Q2); acc = P2;
    } else if (op == 2) acc = acc.dbl();
    else acc = acc.mul(k, 8);
    G1Affine o = acc.to_affine(); memcpy(r, &o, 96);
}
void hc_g2_op(int op, const uint32_t* a, const uint32_t* b, const uint32_t* k, uint32_t* r) {
    G2Affine p, q; memcpy(&p, a, 192); memcpy(&q, b, 192);
    G2XYZZ acc = G2XYZZ::from_affine(p);
    if (op == 0) acc.add_affine(q);
    else if (op == 1) {
        G2XYZZ P2 = acc.dbl(); P2.add_affine(p.neg());
        G2XYZZ Q = G2XYZZ::from_affine(q); G2XYZZ Q2 = Q.dbl(); Q2.add_affine(q.neg());
        P2.add(Q2); acc = P2;
    } else if (op == 2) acc = acc.dbl();
    else acc = acc.mul(k, 8);
    G2Affine o = acc.to_affine(); memcpy(r, &o, 192);
}
}
