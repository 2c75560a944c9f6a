// search.cu -- K4 / K4' / K5: batched lambda-aware search
//   score_i = tau * cos(q, x_i) + (1 - tau) / (1 + |lambda_q - lambda_i|)        (TAUMODE.md:33)
// top-`topk` per query by (score desc, index asc)  (replaces ArrowSpace::search_lambda_aware,
// call site /root/reference/src/lib.rs:173).
//
// Two stages, so that the answer equals the one computed with left-to-right f64 dot products:
//   stage 1  candidate generation.  Scores with tensor-core dot products (different summation order,
//            |s~ - s| <= DELTA), per query the best LIST candidates:
//              search_gemm_kernel  (nq > 8)   FP64 DMMA 128x128 tiles fed by TMA, fused scoring +
//                                             warp-local threshold/top-LIST epilogue in shared memory.
//                                             Bound: FP64 tensor pipe, 2*nq*n*f FLOP.
//              search_gemv_kernel  (nq <= 8)  the reference's one-query-per-call shape.  Bound: HBM,
//                                             8*n*f bytes per pass.
//   stage 2  rescore_kernel.  The LIST candidates of a query are re-scored in the reference order
//            (sequential dot, the oracle's exact expression), sorted, top-k emitted.  The candidate set is
//            complete iff  s~(LIST) < s~(k) - 2 DELTA  (or the shard has <= LIST items); otherwise the
//            query takes the exact full scan (exact_scan_kernel + exact_select_kernel).
//   K5       topk_merge_kernel: merge of per-shard results (cross-GPU), (score desc, index asc).
#include "gemm_topk.cuh"

#include <math.h>

namespace {

using namespace asp_gemm;

// ============================================================================ stage 1: GEMV (nq <= 8)

constexpr int GV_MAXQ = 8;
constexpr int GV_WARPS = 8;

template <int LIST, int FPL>
__global__ void __launch_bounds__(GV_WARPS * 32)
search_gemv_kernel(const double *__restrict__ q, int qpitch, int nq, const double *__restrict__ items, int64_t n_local,
                   int f, int pitch, const double *__restrict__ inv_nx, const double *__restrict__ lam_x,
                   const double *__restrict__ inv_nq, const double *__restrict__ lam_q, double tau,
                   double *__restrict__ cand_score, int32_t *__restrict__ cand_idx)
{
    constexpr int CAP = 2 * LIST;
    constexpr int NPL = CAP / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *qs = reinterpret_cast<double *>(smem_raw);                              // nq * FPL * 32
    double *w_sc = qs + (size_t)GV_MAXQ * FPL * 32;                                 // [warp][q][CAP]
    int32_t *w_ix = reinterpret_cast<int32_t *>(w_sc + GV_WARPS * GV_MAXQ * CAP);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i < nq * FPL * 32; i += blockDim.x) {
        const int qi = i / (FPL * 32), ff = i % (FPL * 32);
        qs[i] = (ff < f) ? q[(size_t)qi * qpitch + ff] : 0.0;
    }
    __syncthreads();

    int cnt[GV_MAXQ];
    double theta[GV_MAXQ];
#pragma unroll
    for (int qi = 0; qi < GV_MAXQ; ++qi) { cnt[qi] = 0; theta[qi] = -INFINITY; }
    const double beta = 1.0 - tau;
    double *my_sc = w_sc + (size_t)warp * GV_MAXQ * CAP;
    int32_t *my_ix = w_ix + (size_t)warp * GV_MAXQ * CAP;

    auto compact = [&](int qi) {
        Cand e[NPL];
#pragma unroll
        for (int t = 0; t < NPL; ++t) {
            const int i = lane + 32 * t;
            if (i < cnt[qi]) { e[t].s = my_sc[qi * CAP + i]; e[t].i = my_ix[qi * CAP + i]; }
            else e[t] = asp::cand_empty();
        }
        asp::warp_sort_best_first<NPL>(e, lane);
#pragma unroll
        for (int t = 0; t < NPL; ++t) {
            const int i = lane + 32 * t;
            if (i < LIST) { my_sc[qi * CAP + i] = e[t].s; my_ix[qi * CAP + i] = e[t].i; }
        }
        const double last = __shfl_sync(0xffffffffu, e[(LIST - 1) / 32].s, (LIST - 1) & 31);
        theta[qi] = (cnt[qi] >= LIST) ? last : -INFINITY;
        cnt[qi] = cnt[qi] < LIST ? cnt[qi] : LIST;
        __syncwarp();
    };

    const int64_t gw = (int64_t)blockIdx.x * GV_WARPS + warp, nw = (int64_t)gridDim.x * GV_WARPS;
    for (int64_t item = gw; item < n_local; item += nw) {
        const double *row = items + item * pitch;
        double xv[FPL];
#pragma unroll
        for (int j = 0; j < FPL; ++j) {
            const int ff = lane + 32 * j;
            xv[j] = (ff < f) ? asp::ld_nc_f64(row + ff) : 0.0;
        }
        const double inx = inv_nx[item], lmx = lam_x[item];
#pragma unroll
        for (int qi = 0; qi < GV_MAXQ; ++qi) {
            if (qi < nq) {
                double d = 0.0;
#pragma unroll
                for (int j = 0; j < FPL; ++j) d = fma(xv[j], qs[(qi * FPL + j) * 32 + lane], d);
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
                const double cs = tau * inv_nq[qi] * d * inx;
                if (cs + beta > theta[qi]) {
                    const double sc = cs + beta / (1.0 + fabs(lam_q[qi] - lmx));
                    if (sc > theta[qi]) {                       // warp-uniform branch
                        if (lane == 0) { my_sc[qi * CAP + cnt[qi]] = sc; my_ix[qi * CAP + cnt[qi]] = (int32_t)item; }
                        cnt[qi]++;
                        __syncwarp();
                        if (cnt[qi] == CAP) compact(qi);
                    }
                }
            }
        }
    }
    // per-warp lists -> per-block list: warp qi merges query qi's GV_WARPS lists
#pragma unroll
    for (int qi = 0; qi < GV_MAXQ; ++qi)
        if (qi < nq) compact(qi);
    __shared__ int s_cnt[GV_WARPS][GV_MAXQ];
    if (lane == 0)
        for (int qi = 0; qi < GV_MAXQ; ++qi) s_cnt[warp][qi] = cnt[qi];
    __syncthreads();
    for (int qi = warp; qi < nq; qi += GV_WARPS) {
        Cand best[NPL];
#pragma unroll
        for (int t = 0; t < NPL; ++t) best[t] = asp::cand_empty();
        for (int w = 0; w < GV_WARPS; ++w) {
            // slots [LIST, CAP) of the merge buffer <- list of warp w; slots [0, LIST) keep the running best
#pragma unroll
            for (int t = 0; t < NPL; ++t) {
                const int i = lane + 32 * t;
                if (i >= LIST) {
                    const int j = i - LIST;
                    if (j < s_cnt[w][qi]) {
                        best[t].s = w_sc[((size_t)w * GV_MAXQ + qi) * CAP + j];
                        best[t].i = w_ix[((size_t)w * GV_MAXQ + qi) * CAP + j];
                    } else best[t] = asp::cand_empty();
                }
            }
            asp::warp_sort_best_first<NPL>(best, lane);
        }
#pragma unroll
        for (int t = 0; t < NPL; ++t) {
            const int i = lane + 32 * t;
            if (i < LIST) {
                const size_t o = ((size_t)qi * gridDim.x + blockIdx.x) * LIST + i;
                cand_score[o] = best[t].s;
                cand_idx[o] = (best[t].s == -INFINITY && best[t].i == 0x7fffffff) ? -1 : best[t].i;
            }
        }
    }
}

__global__ void reciprocal_kernel(const double *__restrict__ x, int64_t n, double *__restrict__ out)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (x[i] > 0.0) ? 1.0 / x[i] : 0.0;
}

// ============================================================================ stage 2: rescore

// the oracle's score expression (oracle.c orc_scores), no contraction
__device__ __forceinline__ double exact_score(double dot, double nq, double nx, double tau, double lq, double lx)
{
    const double den = __dmul_rn(nq, nx);
    const double c = (den == 0.0) ? 0.0 : __ddiv_rn(dot, den);
    const double prox = __ddiv_rn(1.0, __dadd_rn(1.0, fabs(__dsub_rn(lq, lx))));
    return __dadd_rn(__dmul_rn(tau, c), __dmul_rn(__dsub_rn(1.0, tau), prox));
}

constexpr int RS_WARPS = 4;

template <int LIST>
__global__ void __launch_bounds__(RS_WARPS * 32)
rescore_kernel(const double *__restrict__ q, int qpitch, int64_t nq, const double *__restrict__ items, int64_t n_local,
               int f, int pitch, int64_t row0, const double *__restrict__ norm_x, const double *__restrict__ lam_x,
               const double *__restrict__ norm_q, const double *__restrict__ lam_q, double tau, int topk, int nparts,
               const double *__restrict__ cand_score, const int32_t *__restrict__ cand_idx, double delta,
               int64_t *__restrict__ out_idx, double *__restrict__ out_score, int32_t *slow_list, int32_t *slow_count)
{
    constexpr int CAP = 2 * LIST;
    constexpr int NPL = CAP / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *qs_all = reinterpret_cast<double *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *qs = qs_all + (size_t)warp * f;
    const int64_t qi = (int64_t)blockIdx.x * RS_WARPS + warp;
    if (qi >= nq) return;

    for (int j = lane; j < f; j += 32) qs[j] = q[qi * qpitch + j];
    __syncwarp();

    // merge the partial lists: running best in slots [0, LIST), incoming in [LIST, CAP)
    Cand best[NPL];
#pragma unroll
    for (int t = 0; t < NPL; ++t) best[t] = asp::cand_empty();
    for (int p = 0; p < nparts; ++p) {
        const size_t base = ((size_t)qi * nparts + p) * LIST;
#pragma unroll
        for (int t = 0; t < NPL; ++t) {
            const int i = lane + 32 * t;
            if (i >= LIST) {
                const int32_t ci = cand_idx[base + i - LIST];
                if (ci >= 0) { best[t].s = cand_score[base + i - LIST]; best[t].i = ci; }
                else best[t] = asp::cand_empty();
            }
        }
        asp::warp_sort_best_first<NPL>(best, lane);
    }
    // element e lives in lane e%32 slot e/32.  Completeness test on the approximate scores.
    const int kk = topk < n_local ? topk : (int)n_local;
    const double a_k = __shfl_sync(0xffffffffu, best[(kk - 1) / 32].s, (kk - 1) & 31);
    const double a_L = __shfl_sync(0xffffffffu, best[(LIST - 1) / 32].s, (LIST - 1) & 31);
    const bool complete = (n_local <= LIST) || (a_L == -INFINITY) || (a_L < a_k - 2.0 * delta);
    if (!complete || kk > LIST) {
        if (lane == 0) slow_list[atomicAdd(slow_count, 1)] = (int32_t)qi;
        return;
    }
    // exact rescoring of the LIST candidates (one lane each), reference order
    const double nqv = norm_q[qi], lqv = lam_q[qi];
#pragma unroll
    for (int t = 0; t < NPL; ++t) {
        const int i = lane + 32 * t;
        if (i < LIST && best[t].i != 0x7fffffff) {
            const int64_t it = best[t].i;
            const double d = seq_dot_row(qs, items + it * pitch, f);
            best[t].s = exact_score(d, nqv, norm_x[it], tau, lqv, lam_x[it]);
        } else best[t] = asp::cand_empty();
    }
    asp::warp_sort_best_first<NPL>(best, lane);
#pragma unroll
    for (int t = 0; t < NPL; ++t) {
        const int i = lane + 32 * t;
        if (i < topk) {
            const bool ok = (i < kk) && best[t].i != 0x7fffffff;
            out_idx[qi * topk + i] = ok ? row0 + best[t].i : -1;
            out_score[qi * topk + i] = ok ? best[t].s : NAN;
        }
    }
}

// ============================================================================ slow path: exact full scan
// Queries whose candidate set could not be proven complete (ties around the k-th score, emission overflow, topk >= 32) are
// answered from the reference-order score of EVERY item.  All such queries of a batch go through two launches:
//   exact_scan_topk_kernel  block (x, y) = item range x of XS_QPB queries: a thread owns an item row at a time and carries the
//                           XS_QPB ordered dot products together (the row is read once for all of them); the scores of a
//                           chunk land in shared memory, one warp per query folds them into that query's running top-k
//                           (k rounds of "best entry strictly after the previous winner": exact ties by index);
//   exact_merge_kernel      one block per query: the same k rounds over the ranges' top-k lists.
constexpr int XS_THREADS = 256, XS_CHUNK = 1024, XS_TOPK_MAX = 1024;   // XS_QPB (template): 4 queries per block, 1 for very wide rows

// (score desc, index asc): does (s, i) come strictly after (ps, pi), and before (bs, bi)?
__device__ __forceinline__ bool xs_after(double s, int64_t i, double ps, int64_t pi) { return (s < ps) || (s == ps && i > pi); }
__device__ __forceinline__ bool xs_before(double s, int64_t i, double bs, int64_t bi) { return (s > bs) || (s == bs && i < bi); }

template <int XS_QPB>
__global__ void __launch_bounds__(XS_THREADS)
exact_scan_topk_kernel(const double *__restrict__ q, int qpitch, const int32_t *__restrict__ slow_list, int nslow,
                       const double *__restrict__ items, int64_t n_local, int f, int pitch,
                       const double *__restrict__ norm_x, const double *__restrict__ lam_x,
                       const double *__restrict__ norm_q, const double *__restrict__ lam_q, double tau, int topk,
                       double *__restrict__ part_score, int64_t *__restrict__ part_idx)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *qs = reinterpret_cast<double *>(smem_raw);                        // [XS_QPB][f]
    double *sc = qs + (size_t)XS_QPB * f;                                    // [XS_QPB][2 topk + XS_CHUNK]: top | chunk | new top
    int64_t *ix = reinterpret_cast<int64_t *>(sc + (size_t)XS_QPB * (2 * topk + XS_CHUNK));
    const int stride = 2 * topk + XS_CHUNK;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.y * XS_QPB;
    int64_t qi[XS_QPB];
    double nqv[XS_QPB], lqv[XS_QPB];
#pragma unroll
    for (int u = 0; u < XS_QPB; ++u) {
        qi[u] = (q0 + u < nslow) ? slow_list[q0 + u] : -1;
        nqv[u] = qi[u] >= 0 ? norm_q[qi[u]] : 1.0;
        lqv[u] = qi[u] >= 0 ? lam_q[qi[u]] : 0.0;
    }
    for (int j = threadIdx.x; j < XS_QPB * f; j += XS_THREADS) {
        const int u = j / f, t = j - u * f;
        const int64_t src = (q0 + u < nslow) ? slow_list[q0 + u] : -1;
        qs[j] = src >= 0 ? q[src * qpitch + t] : 0.0;
    }
    for (int j = threadIdx.x; j < XS_QPB * topk; j += XS_THREADS) {
        const int u = j / topk, t = j - u * topk;
        sc[u * stride + t] = -INFINITY;
        ix[u * stride + t] = INT64_MAX;
    }
    __syncthreads();
    const int64_t per = (n_local + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = blockIdx.x * per, r1 = (r0 + per < n_local) ? r0 + per : n_local;
    for (int64_t c0 = r0; c0 < r1; c0 += XS_CHUNK) {
        const int len = (int)((r1 - c0 < XS_CHUNK) ? r1 - c0 : XS_CHUNK);
        for (int e = threadIdx.x; e < len; e += XS_THREADS) {
            const int64_t it = c0 + e;
            const double *row = items + it * pitch;
            double d[XS_QPB];
#pragma unroll
            for (int u = 0; u < XS_QPB; ++u) d[u] = 0.0;
            int j = 0;
            for (; j + 2 <= f; j += 2) {                                      // one load of the row for the XS_QPB ordered sums
                const double2 xv = *reinterpret_cast<const double2 *>(row + j);
#pragma unroll
                for (int u = 0; u < XS_QPB; ++u) {
                    d[u] = __dadd_rn(d[u], __dmul_rn(qs[u * f + j], xv.x));
                    d[u] = __dadd_rn(d[u], __dmul_rn(qs[u * f + j + 1], xv.y));
                }
            }
            for (; j < f; ++j)
#pragma unroll
                for (int u = 0; u < XS_QPB; ++u) d[u] = __dadd_rn(d[u], __dmul_rn(qs[u * f + j], row[j]));
            const double nx = norm_x[it], lx = lam_x[it];
#pragma unroll
            for (int u = 0; u < XS_QPB; ++u) {
                sc[u * stride + topk + e] = exact_score(d[u], nqv[u], nx, tau, lqv[u], lx);
                ix[u * stride + topk + e] = it;
            }
        }
        __syncthreads();
        if (warp < XS_QPB && qi[warp] >= 0) {                                 // warp u folds the chunk into query u's top-k
            double *s_u = sc + warp * stride;
            int64_t *i_u = ix + warp * stride;
            const int total = topk + len;
            double ps = INFINITY;
            int64_t pi = -1;
            for (int r = 0; r < topk; ++r) {
                double bs = -INFINITY;
                int64_t bi = INT64_MAX;
                for (int e = lane; e < total; e += 32) {
                    const double v = s_u[e];
                    const int64_t vi = i_u[e];
                    if (vi != INT64_MAX && xs_after(v, vi, ps, pi) && xs_before(v, vi, bs, bi)) { bs = v; bi = vi; }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const double os = __shfl_xor_sync(0xffffffffu, bs, off);
                    const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, off);
                    if (xs_before(os, oi, bs, bi)) { bs = os; bi = oi; }
                }
                if (lane == 0) { s_u[topk + XS_CHUNK + r] = bs; i_u[topk + XS_CHUNK + r] = bi; }
                if (bi == INT64_MAX) { ps = -INFINITY; pi = INT64_MAX; } else { ps = bs; pi = bi; }
            }
            __syncwarp();
            for (int r = lane; r < topk; r += 32) { s_u[r] = s_u[topk + XS_CHUNK + r]; i_u[r] = i_u[topk + XS_CHUNK + r]; }
        }
        __syncthreads();
    }
    for (int j = threadIdx.x; j < XS_QPB * topk; j += XS_THREADS) {
        const int u = j / topk, t = j - u * topk;
        if (q0 + u < nslow) {
            const size_t o = ((size_t)(q0 + u) * gridDim.x + blockIdx.x) * topk + t;
            part_score[o] = sc[u * stride + t];
            part_idx[o] = ix[u * stride + t];
        }
    }
}

// one block per slow query: top-k of its nparts x topk range winners
__global__ void __launch_bounds__(XS_THREADS)
exact_merge_kernel(const double *__restrict__ part_score, const int64_t *__restrict__ part_idx, int nparts, int topk,
                   int64_t row0, const int32_t *__restrict__ slow_list, int64_t *__restrict__ out_idx, double *__restrict__ out_score)
{
    __shared__ double s_s[XS_THREADS / 32];
    __shared__ int64_t s_i[XS_THREADS / 32];
    __shared__ double prev_s;
    __shared__ int64_t prev_i;
    const int64_t qi = slow_list[blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t base = (size_t)blockIdx.x * nparts * topk;
    const int total = nparts * topk;
    if (threadIdx.x == 0) { prev_s = INFINITY; prev_i = -1; }
    __syncthreads();
    for (int r = 0; r < topk; ++r) {
        const double ps = prev_s;
        const int64_t pi = prev_i;
        double bs = -INFINITY;
        int64_t bi = INT64_MAX;
        for (int e = threadIdx.x; e < total; e += XS_THREADS) {
            const double v = part_score[base + e];
            const int64_t vi = part_idx[base + e];
            if (vi != INT64_MAX && xs_after(v, vi, ps, pi) && xs_before(v, vi, bs, bi)) { bs = v; bi = vi; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double os = __shfl_xor_sync(0xffffffffu, bs, off);
            const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (xs_before(os, oi, bs, bi)) { bs = os; bi = oi; }
        }
        if (lane == 0) { s_s[warp] = bs; s_i[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            bs = (lane < XS_THREADS / 32) ? s_s[lane] : -INFINITY;
            bi = (lane < XS_THREADS / 32) ? s_i[lane] : INT64_MAX;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double os = __shfl_xor_sync(0xffffffffu, bs, off);
                const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (xs_before(os, oi, bs, bi)) { bs = os; bi = oi; }
            }
            if (lane == 0) {
                const bool ok = (bi != INT64_MAX);
                out_idx[qi * topk + r] = ok ? row0 + bi : -1;
                out_score[qi * topk + r] = ok ? bs : NAN;
                prev_s = ok ? bs : -INFINITY;
                prev_i = ok ? bi : INT64_MAX;
            }
        }
        __syncthreads();
    }
}

// ============================================================================ K5: cross-shard merge

struct MKey { double s; int64_t i; };
__device__ __forceinline__ bool mkey_better(const MKey &x, const MKey &y)
{
    return (x.s > y.s) || (x.s == y.s && x.i < y.i);
}

// one block per query; bitonic sort of parts*topk entries in shared memory
__global__ void topk_merge_kernel(const int64_t *__restrict__ idx, const double *__restrict__ score, int parts, int64_t nq,
                                  int topk, int p2, int64_t *__restrict__ out_idx, double *__restrict__ out_score)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    MKey *keys = reinterpret_cast<MKey *>(smem_raw);
    const int64_t qi = blockIdx.x;
    const int total = parts * topk;
    for (int i = threadIdx.x; i < p2; i += blockDim.x) {
        MKey k; k.s = -INFINITY; k.i = INT64_MAX;
        if (i < total) {
            const int p = i / topk, j = i % topk;
            const int64_t id = idx[((size_t)p * nq + qi) * topk + j];
            if (id >= 0) { k.s = score[((size_t)p * nq + qi) * topk + j]; k.i = id; }
        }
        keys[i] = k;
    }
    for (int size = 2; size <= p2; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int i = threadIdx.x; i < p2 / 2; i += blockDim.x) {
                const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
                const bool first = ((lo & size) == 0);
                const MKey a = keys[lo], b = keys[hi];
                if (first ? mkey_better(b, a) : mkey_better(a, b)) { keys[lo] = b; keys[hi] = a; }
            }
        }
    __syncthreads();
    for (int i = threadIdx.x; i < topk; i += blockDim.x) {
        const bool ok = (i < p2) && keys[i].i != INT64_MAX;
        out_idx[qi * topk + i] = ok ? keys[i].i : -1;
        out_score[qi * topk + i] = ok ? keys[i].s : NAN;
    }
}

template <int LIST>
int launch_gemm(const asp_space *s, const CUtensorMap &tmap_q, const double *q_dev, int64_t nq, const double *inv_nq,
                const double *lam_q, double tau, int nchunks, double *cand_score, int32_t *cand_idx)
{
    asp_ctx *ctx = s->ctx;
    constexpr int STAGES = (LIST == 16) ? 4 : 3;
    const size_t smem = (size_t)STAGES * STAGE_DOUBLES_S * 8 + sizeof(ListSmem<LIST>) + 128;
    dim3 grid((unsigned)asp_ceil_div(nq, QT), nchunks);
    if (ctx->use_tma) {
        auto k = search_gemm_kernel<LIST, STAGES, true, 0>;
        ASP_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, MMA_WARPS * 32, smem, ctx->stream>>>(tmap_q, s->tmap_rows, q_dev, s->items, nq, s->n_local, s->fp,
                                                           s->inv_norms, s->lambdas, inv_nq, lam_q, tau, 0.0, nchunks, cand_score,
                                                           cand_idx);
    } else {
        auto k = search_gemm_kernel<LIST, STAGES, false, 0>;
        ASP_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, MMA_WARPS * 32, smem, ctx->stream>>>(tmap_q, s->tmap_rows, q_dev, s->items, nq, s->n_local, s->fp,
                                                     s->inv_norms, s->lambdas, inv_nq, lam_q, tau, 0.0, nchunks, cand_score,
                                                     cand_idx);
    }
    ASP_CUDA(cudaGetLastError());
    ASP_LAUNCHED(ctx);
    return ASP_OK;
}

template <int LIST>
int launch_gemv(const asp_space *s, const double *q_dev, int qpitch, int nq, const double *inv_nq, const double *lam_q,
                double tau, int nblocks, double *cand_score, int32_t *cand_idx)
{
    asp_ctx *ctx = s->ctx;
    constexpr int CAP = 2 * LIST;
    const int f = s->f;
    int fpl = (f + 31) / 32;
    auto smem_for = [&](int FPL) { return (size_t)GV_MAXQ * FPL * 32 * 8 + (size_t)GV_WARPS * GV_MAXQ * CAP * 12; };
#define ASP_GEMV_CASE(FPLV)                                                                                         \
    {                                                                                                               \
        auto k = search_gemv_kernel<LIST, FPLV>;                                                                    \
        const size_t smem = smem_for(FPLV);                                                                         \
        ASP_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                  \
        k<<<nblocks, GV_WARPS * 32, smem, ctx->stream>>>(q_dev, qpitch, nq, s->items, s->n_local, f, s->fp,         \
                                                        s->inv_norms, s->lambdas, inv_nq, lam_q, tau, cand_score,   \
                                                        cand_idx);                                                  \
    }
    if (fpl <= 4) ASP_GEMV_CASE(4)
    else if (fpl <= 12) ASP_GEMV_CASE(12)
    else if (fpl <= 24) ASP_GEMV_CASE(24)
    else if (fpl <= 48) ASP_GEMV_CASE(48)
    else ASP_FAIL(ASP_ERR_UNSUPPORTED, "single-query search supports at most 1536 features (got %d)", f);
#undef ASP_GEMV_CASE
    ASP_CUDA(cudaGetLastError());
    ASP_LAUNCHED(ctx);
    return ASP_OK;
}

}  // namespace

int asp_launch_reciprocal(asp_ctx *ctx, const double *x, int64_t n, double *out)
{
    reciprocal_kernel<<<(unsigned)(asp_ceil_div(n, 256) < 1024 ? asp_ceil_div(n, 256) : 1024), 256, 0, ctx->stream>>>(x, n, out);
    ASP_CUDA(cudaGetLastError());
    ASP_LAUNCHED(ctx);
    return ASP_OK;
}

// exact full scan of the listed queries (completeness test failed / emission buffer full)
int asp_search_slow_path(const asp_space *s, const double *q_dev, int32_t qpitch, const double *lambda_q_dev,
                         const double *qnorm_dev, double tau, int64_t topk, const int32_t *slow_list, int nslow,
                         int64_t *out_idx_dev, double *out_score_dev)
{
    asp_ctx *ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    const int f = s->f;
    if (nslow <= 0) return ASP_OK;
    if (topk > XS_TOPK_MAX) ASP_FAIL(ASP_ERR_UNSUPPORTED, "exact scan supports topk <= %d (got %lld)", XS_TOPK_MAX, (long long)topk);
    auto smem_for = [&](int qpb) { return sizeof(double) * (size_t)qpb * f + (size_t)qpb * (2 * topk + XS_CHUNK) * 16; };
    const int qpb = (smem_for(4) <= 200 * 1024) ? 4 : 1;
    const size_t smem = smem_for(qpb);
    if (smem > 220 * 1024) ASP_FAIL(ASP_ERR_UNSUPPORTED, "exact scan: %d features x topk %lld do not fit in shared memory", f, (long long)topk);
    // item ranges: enough blocks to fill the device a few times over per query group, at least one chunk each
    int64_t nparts = std::min<int64_t>(asp_ceil_div(s->n_local, XS_CHUNK), (int64_t)ctx->num_sms * 2);
    if (nparts < 1) nparts = 1;
    const int64_t ngroups = asp_ceil_div(nslow, qpb);
    double *part_score = nullptr;
    int64_t *part_idx = nullptr;
    if (qpb == 4) ASP_CUDA(cudaFuncSetAttribute(exact_scan_topk_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else ASP_CUDA(cudaFuncSetAttribute(exact_scan_topk_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // groups of queries per launch: bounded scratch (nslow can be the whole batch when topk >= 32) and grid.y <= 65535
    const int64_t max_groups = std::max<int64_t>(1, std::min<int64_t>(65535, (int64_t)(1u << 28) / (nparts * topk * 16 * qpb)));
    ASP_CUDA(cudaMallocAsync(&part_score, sizeof(double) * (size_t)std::min(ngroups, max_groups) * qpb * nparts * topk, st));
    ASP_CUDA(cudaMallocAsync(&part_idx, sizeof(int64_t) * (size_t)std::min(ngroups, max_groups) * qpb * nparts * topk, st));
    for (int64_t g0 = 0; g0 < ngroups; g0 += max_groups) {
        const int64_t ng = std::min(max_groups, ngroups - g0);
        const int first = (int)(g0 * qpb), count = (int)std::min<int64_t>(ng * qpb, nslow - first);
        const dim3 grid((unsigned)nparts, (unsigned)ng);
        if (qpb == 4)
            exact_scan_topk_kernel<4><<<grid, XS_THREADS, smem, st>>>(q_dev, qpitch, slow_list + first, count, s->items, s->n_local, f,
                                                                      s->fp, s->norms, s->lambdas, qnorm_dev, lambda_q_dev, tau,
                                                                      (int)topk, part_score, part_idx);
        else
            exact_scan_topk_kernel<1><<<grid, XS_THREADS, smem, st>>>(q_dev, qpitch, slow_list + first, count, s->items, s->n_local, f,
                                                                      s->fp, s->norms, s->lambdas, qnorm_dev, lambda_q_dev, tau,
                                                                      (int)topk, part_score, part_idx);
        ASP_CUDA(cudaGetLastError()); ASP_LAUNCHED(ctx);
        exact_merge_kernel<<<(unsigned)count, XS_THREADS, 0, st>>>(part_score, part_idx, (int)nparts, (int)topk, s->row0,
                                                                  slow_list + first, out_idx_dev, out_score_dev);
        ASP_CUDA(cudaGetLastError()); ASP_LAUNCHED(ctx);
    }
    ASP_CUDA(cudaFreeAsync(part_score, st));
    ASP_CUDA(cudaFreeAsync(part_idx, st));
    return ASP_OK;
}

int asp_search_impl(const asp_space *s, const asp_graph *g, const double *q_dev, int64_t nq, int32_t qpitch,
                    const double *lambda_q_dev, const double *qnorm_dev, double tau, int64_t topk, int64_t *out_idx_dev,
                    double *out_score_dev)
{
    (void)g;
    asp_ctx *ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    if (nq == 0 || topk == 0) return ASP_OK;
    const int f = s->f;
    const int LISTSEL = (topk <= 12) ? 16 : 32;
    // rounding band of one approximate score: dot products of f terms in two orders, both norms,
    // the blend: (4 f + 64) u  scaled by |tau| + |1 - tau|
    const double u = 1.1102230246251565e-16;
    const double delta = (4.0 * f + 64.0) * u * (fabs(tau) + fabs(1.0 - tau) + 1.0);

    // 1/norm of the queries
    double *inv_nq = nullptr;
    ASP_CUDA(cudaMallocAsync(&inv_nq, sizeof(double) * nq, st));
    ASP_CHECK(asp_launch_reciprocal(ctx, qnorm_dev, nq, inv_nq));

    int32_t *slow_list = nullptr, *slow_count = nullptr;
    ASP_CUDA(cudaMallocAsync(&slow_list, sizeof(int32_t) * (nq + 1), st));
    ASP_CUDA(cudaMallocAsync(&slow_count, sizeof(int32_t), st));
    ASP_CUDA(cudaMemsetAsync(slow_count, 0, sizeof(int32_t), st));

    int nparts = 0;
    double *cand_score = nullptr;
    int32_t *cand_idx = nullptr;
    ASP_CUDA(cudaEventRecord(ctx->ev0, st));
    if (nq <= GV_MAXQ && f <= 1536) {            // the GEMV kernel keeps a query in registers: wider vectors take the DMMA tile kernel
        int64_t want = asp_ceil_div(s->n_local, GV_WARPS * 16);
        nparts = (int)(want < ctx->num_sms * 2 ? (want > 0 ? want : 1) : ctx->num_sms * 2);
        ASP_CUDA(cudaMallocAsync(&cand_score, sizeof(double) * (size_t)nq * nparts * LISTSEL, st));
        ASP_CUDA(cudaMallocAsync(&cand_idx, sizeof(int32_t) * (size_t)nq * nparts * LISTSEL, st));
        if (LISTSEL == 16) ASP_CHECK(launch_gemv<16>(s, q_dev, qpitch, (int)nq, inv_nq, lambda_q_dev, tau, nparts, cand_score, cand_idx));
        else ASP_CHECK(launch_gemv<32>(s, q_dev, qpitch, (int)nq, inv_nq, lambda_q_dev, tau, nparts, cand_score, cand_idx));
    } else {
        const int64_t tiles_total = asp_ceil_div(s->n_local, IT);
        const int64_t qblocks = asp_ceil_div(nq, QT);
        // CTAs = qblocks x chunks; pick the chunk count whose CTA total fills whole waves of SMs
        // (1 CTA per SM): smallest wave count w with >= 97 % of w * num_sms CTAs, else the best seen.
        int64_t best_chunks = 1;
        double best_eff = 0.0;
        for (int w = 1; w <= 16; ++w) {
            int64_t c = ((int64_t)w * ctx->num_sms) / qblocks;
            if (c < 1) continue;
            if (c > tiles_total) c = tiles_total;
            const int64_t ctas = c * qblocks;
            const double eff = (double)ctas / (double)(asp_ceil_div(ctas, ctx->num_sms) * ctx->num_sms);
            if (eff > best_eff + 1e-9) { best_eff = eff; best_chunks = c; }
            if (eff >= 0.97 || c == tiles_total) break;
        }
        nparts = (int)best_chunks;
        CUtensorMap tmap_q;
        ASP_CHECK(asp_make_items_tmap(&tmap_q, q_dev, nq, qpitch, QT, KSTEP / 4));
        ASP_CUDA(cudaMallocAsync(&cand_score, sizeof(double) * (size_t)nq * nparts * LISTSEL, st));
        ASP_CUDA(cudaMallocAsync(&cand_idx, sizeof(int32_t) * (size_t)nq * nparts * LISTSEL, st));
        if (LISTSEL == 16) ASP_CHECK(launch_gemm<16>(s, tmap_q, q_dev, nq, inv_nq, lambda_q_dev, tau, nparts, cand_score, cand_idx));
        else ASP_CHECK(launch_gemm<32>(s, tmap_q, q_dev, nq, inv_nq, lambda_q_dev, tau, nparts, cand_score, cand_idx));
    }
    ASP_CUDA(cudaEventRecord(ctx->ev1, st));

    {
        const size_t smem = (size_t)RS_WARPS * f * 8;
        const unsigned grid = (unsigned)asp_ceil_div(nq, RS_WARPS);
        if (LISTSEL == 16) {
            ASP_CUDA(cudaFuncSetAttribute(rescore_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            rescore_kernel<16><<<grid, RS_WARPS * 32, smem, st>>>(q_dev, qpitch, nq, s->items, s->n_local, f, s->fp, s->row0,
                                                                 s->norms, s->lambdas, qnorm_dev, lambda_q_dev, tau, (int)topk,
                                                                 nparts, cand_score, cand_idx, delta, out_idx_dev,
                                                                 out_score_dev, slow_list, slow_count);
        } else {
            ASP_CUDA(cudaFuncSetAttribute(rescore_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            rescore_kernel<32><<<grid, RS_WARPS * 32, smem, st>>>(q_dev, qpitch, nq, s->items, s->n_local, f, s->fp, s->row0,
                                                                 s->norms, s->lambdas, qnorm_dev, lambda_q_dev, tau, (int)topk,
                                                                 nparts, cand_score, cand_idx, delta, out_idx_dev,
                                                                 out_score_dev, slow_list, slow_count);
        }
        ASP_CUDA(cudaGetLastError());
        ASP_LAUNCHED(ctx);
    }

    int32_t nslow = 0;
    ASP_CUDA(cudaMemcpyAsync(&nslow, slow_count, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    ASP_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->stats["search_stage1_ms"] = ms;
    ctx->stats["search_slow_queries"] = nslow;
    ctx->stats["search_stage1_is_tc"] = 0.0;
    if (nslow > 0)
        ASP_CHECK(asp_search_slow_path(s, q_dev, qpitch, lambda_q_dev, qnorm_dev, tau, topk, slow_list, nslow, out_idx_dev,
                                       out_score_dev));
    ASP_CUDA(cudaFreeAsync(cand_score, st));
    ASP_CUDA(cudaFreeAsync(cand_idx, st));
    ASP_CUDA(cudaFreeAsync(slow_list, st));
    ASP_CUDA(cudaFreeAsync(slow_count, st));
    ASP_CUDA(cudaFreeAsync(inv_nq, st));
    return ASP_OK;
}

int asp_topk_merge_impl(asp_ctx *ctx, const int64_t *idx_dev, const double *score_dev, int parts, int64_t nq, int64_t topk,
                        int64_t *out_idx_dev, double *out_score_dev)
{
    if (nq == 0 || topk == 0) return ASP_OK;
    int p2 = 1;
    while (p2 < parts * topk) p2 <<= 1;
    const size_t smem = (size_t)p2 * sizeof(MKey);
    if (smem > 200 * 1024) ASP_FAIL(ASP_ERR_UNSUPPORTED, "topk merge of %d x %lld entries does not fit in shared memory", parts, (long long)topk);
    ASP_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    topk_merge_kernel<<<(unsigned)nq, 128, smem, ctx->stream>>>(idx_dev, score_dev, parts, nq, (int)topk, p2, out_idx_dev,
                                                                out_score_dev);
    ASP_CUDA(cudaGetLastError());
    ASP_LAUNCHED(ctx);
    return ASP_OK;
}
