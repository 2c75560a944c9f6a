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

// Sorts 32*NPL elements held as e[t] = element (lane + 32 t), best first in element order.  Branch free: every
// compare-exchange is a predicate and two selects (distinct elements never compare equal; two empties are identical).
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
                const int ts = stride >> 5;
#pragma unroll
                for (int t = 0; t < NPL; ++t) {
                    if ((t & ts) == 0) {
                        const int i = lane + 32 * t;
                        const bool first_block = ((i & size) == 0);      // "ascending" = best first
                        const Cand lo = e[t], hi = e[t | ts];
                        const bool swap = (cand_better(hi, lo) == first_block);
                        e[t].s = swap ? hi.s : lo.s;
                        e[t].i = swap ? hi.i : lo.i;
                        e[t | ts].s = swap ? lo.s : hi.s;
                        e[t | ts].i = swap ? lo.i : hi.i;
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
                    const bool take = (cand_better(o, e[t]) == keep_best) && !(o.s == e[t].s && o.i == e[t].i);
                    e[t].s = take ? o.s : e[t].s;
                    e[t].i = take ? o.i : e[t].i;
                }
            }
        }
    }
}

// Descending sort of 64 floats (2 per lane, element = lane + 32 t): max/min exchanges only.
__device__ __forceinline__ void warp_sort_desc_f32x2(float (&v)[2], int lane)
{
#pragma unroll
    for (int size = 2; size <= 64; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride == 32) {
                const float a = fmaxf(v[0], v[1]), b = fminf(v[0], v[1]);   // size == 64: one descending block
                v[0] = a;
                v[1] = b;
            } else {
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const int i = lane + 32 * t;
                    const float o = __shfl_xor_sync(0xffffffffu, v[t], stride);
                    const bool first_block = ((i & size) == 0);
                    const bool lower = ((i & stride) == 0);
                    v[t] = (lower == first_block) ? fmaxf(v[t], o) : fminf(v[t], o);
                }
            }
        }
    }
}

}  // namespace asp
