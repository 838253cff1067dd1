// Pair lists of the round-based bucket accumulation (msm_pairs.cuh): the rule that turns ONE bucket
// slice into its pairs of every round.  Plain C++ (host + device): used by pair_build_kernel
// (msm_sort_kernels.cuh), by the host planner, and by tests/host_check.
#pragma once
#include <stdint.h>

#include "field.cuh"

namespace bmpc {

#define BMPC_PAIR_PAD 0xffffffffu
#define BMPC_PAIR_NONE 0xffffffffu
#define BMPC_PAIR_MAX_ROUNDS 12          // slices of at most 2^12 entries

struct PairIdx { uint32_t a, b; };       // == uint2 on the device

// elements of a slice of c0 (even) entries that are left after r rounds, and its pairs in round r
BMPC_HD uint32_t pair_elems(uint32_t c0, uint32_t r) { return (uint32_t)(((uint64_t)c0 + ((1ull << r) - 1)) >> r); }
BMPC_HD uint32_t pair_count(uint32_t c0, uint32_t r) { return pair_elems(c0, r) >> 1; }

// Lists of ONE slice for rounds 1 .. R-1.  start/c0: the slice's (even) first position and padded
// length in sorted'; pairoff[r]: first list position of this slice in round r; out_base[r]: pool
// index of round r's first result; lists[r]: round r's list.  Returns the pool index of the slice's
// sum (BMPC_PAIR_NONE for an empty slice).
BMPC_HD uint32_t pair_build_task(uint32_t start, uint32_t c0, uint32_t R, const uint32_t* pairoff,
                                 const uint32_t* out_base, PairIdx* const* lists) {
    if (c0 == 0) return BMPC_PAIR_NONE;
    uint32_t cur_base = out_base[0] + (start >> 1), cur_cnt = c0 >> 1, carry = BMPC_PAIR_NONE;
    for (uint32_t r = 1; r < R; r++) {
        const uint32_t total = cur_cnt + (carry != BMPC_PAIR_NONE ? 1u : 0u);
        if (total <= 1) break;
        PairIdx* L = lists[r] + pairoff[r];
        const uint32_t full = cur_cnt >> 1;
        for (uint32_t i = 0; i < full; i++) {
            PairIdx e;
            e.a = cur_base + 2 * i;
            e.b = cur_base + 2 * i + 1;
            L[i] = e;
        }
        uint32_t np = full;
        if (cur_cnt & 1u) {
            const uint32_t last = cur_base + cur_cnt - 1;
            if (carry != BMPC_PAIR_NONE) {
                PairIdx e;
                e.a = last;
                e.b = carry;
                L[full] = e;
                np = full + 1;
                carry = BMPC_PAIR_NONE;
            } else {
                carry = last;
            }
        }
        cur_base = out_base[r] + pairoff[r];
        cur_cnt = np;
    }
    return cur_cnt ? cur_base : carry;      // one element left: the last result, or the carried one
}

}  // namespace bmpc
