// warp_sort.cuh -- warp-level bitonic sort of (score, index) candidates through shuffles.
// Order: better first = larger score, ties -> smaller index (SURVEY.md Appendix A9 tie rule).
#pragma once
#include <stdint.h>

namespace asp {

struct Cand {
    double s;
    int32_t i;
};

__device__ __forceinline__ bool cand_better(const Cand &x, const Cand &y)
{
    return (x.s > y.s) || (x.s == y.s && x.i < y.i);
}

__device__ __forceinline__ Cand cand_empty()
{
    Cand c;
    c.s = -INFINITY;
    c.i = 0x7fffffff;
    return c;
}

// Sorts 32*NPL elements held as e[t] = element (lane + 32 t), best first in element order.
template <int NPL>
__device__ __forceinline__ void warp_sort_best_first(Cand (&e)[NPL], int lane)
{
    constexpr int N = 32 * NPL;
#pragma unroll
    for (int size = 2; size <= N; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= 32) {
                // partner lives in the same lane, slot t ^ (stride/32)
                constexpr int dummy = 0;
                (void)dummy;
                const int ts = stride >> 5;
#pragma unroll
                for (int t = 0; t < NPL; ++t) {
                    if ((t & ts) == 0) {
                        const int i = lane + 32 * t;
                        const bool first_block = ((i & size) == 0);      // "ascending" = best first
                        Cand &lo = e[t], &hi = e[t | ts];
                        const bool swap = first_block ? cand_better(hi, lo) : cand_better(lo, hi);
                        if (swap) { Cand tmp = lo; lo = hi; hi = tmp; }
                    }
                }
            } else {
#pragma unroll
                for (int t = 0; t < NPL; ++t) {
                    const int i = lane + 32 * t;
                    Cand o;
                    o.s = __shfl_xor_sync(0xffffffffu, e[t].s, stride);
                    o.i = __shfl_xor_sync(0xffffffffu, e[t].i, stride);
                    const bool first_block = ((i & size) == 0);
                    const bool lower = ((i & stride) == 0);
                    const bool keep_best = (lower == first_block);
                    const bool o_better = cand_better(o, e[t]);
                    const bool e_better = cand_better(e[t], o);
                    if (keep_best ? o_better : e_better) e[t] = o;
                }
            }
        }
    }
}

}  // namespace asp
