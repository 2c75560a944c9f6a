// hybrid.cuh -- re-ranking of a cosine shortlist: the second half of asp_search_hybrid_batch.
//
// Replaces the crate's search_lambda_aware_hybrid as called by ArrowSpace.search_hybrid (src/lib.rs:182-219).  Its body is
// not in the reference (PARITY UNPINNED); the restatement H1-H3 is in include/arrowspace_b200.h and oracle/oracle.c.  The
// shortlist itself (H2) is the validated search at tau = 1 with topk = pool (search_tc.cu / search.cu); this file holds H3:
//   hybrid_rescore_kernel  one thread per (query, shortlist slot): the cosine re-evaluated in the oracle's order (left to
//                          right, product rounded then added) and the lambda-aware score in the oracle's expression, so the
//                          scores are the oracle's bit for bit;
//   hybrid_select_kernel   one thread per query: the best topk slots by (score desc, index asc).
// Both are a few hundred steps per query next to a scan of every item for the shortlist; no shared memory, no barriers.
//
// The per-thread bodies are plain functions so that tests/test_hybrid_host.py can compile this header with g++
// (-ffp-contract=off) and walk the thread grid on the CPU against the oracle.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define ASP_HYB_FN __host__ __device__ __forceinline__
#else
#define ASP_HYB_FN static inline
#endif

namespace asp_hybrid {

ASP_HYB_FN double mul_rn(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
ASP_HYB_FN double add_rn(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
ASP_HYB_FN double sub_rn(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
ASP_HYB_FN double div_rn(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}

// H3 for shortlist slot t = qi * pool + j.  pool_idx holds GLOBAL row indices (-1 = padding: fewer items than slots).
// Scores of padding slots are NaN and never selected.
ASP_HYB_FN void rescore_slot(int64_t t, int64_t pool, const double *q, int qpitch, const double *items, int pitch, int f,
                             int64_t row0, const double *norm_x, const double *lam_x, const double *norm_q,
                             const double *lam_q, double tau, const int64_t *pool_idx, double *pool_score)
{
    const int64_t qi = t / pool;
    const int64_t id = pool_idx[t];
    if (id < 0) { pool_score[t] = NAN; return; }
    const int64_t li = id - row0;
    const double *qv = q + (size_t)qi * qpitch;
    const double *xv = items + (size_t)li * pitch;
    double dot = 0.0;
    for (int c = 0; c < f; ++c) dot = add_rn(dot, mul_rn(qv[c], xv[c]));          // oracle.c seq_dot
    const double den = mul_rn(norm_q[qi], norm_x[li]);
    const double cs = (den == 0.0) ? 0.0 : div_rn(dot, den);                        // README.md:69 arithmetic
    const double prox = div_rn(1.0, add_rn(1.0, fabs(sub_rn(lam_q[qi], lam_x[li]))));
    pool_score[t] = add_rn(mul_rn(tau, cs), mul_rn(sub_rn(1.0, tau), prox));       // TAUMODE.md:33
}

// The best topk of query qi's shortlist by (score desc, index asc); pool_idx is scratch (taken slots are overwritten with -1).
// Output rows are padded with -1 / NaN like asp_search_batch.
ASP_HYB_FN void select_query(int64_t qi, int64_t pool, int64_t topk, int64_t *pool_idx, const double *pool_score,
                             int64_t *out_idx, double *out_score)
{
    int64_t *ids = pool_idx + (size_t)qi * pool;
    const double *sc = pool_score + (size_t)qi * pool;
    for (int64_t r = 0; r < topk; ++r) {
        int64_t best = -1;
        for (int64_t j = 0; j < pool; ++j) {
            if (ids[j] < 0) continue;
            if (best < 0 || sc[j] > sc[best] || (sc[j] == sc[best] && ids[j] < ids[best])) best = j;
        }
        if (best < 0) {
            out_idx[qi * topk + r] = -1;
            out_score[qi * topk + r] = NAN;
        } else {
            out_idx[qi * topk + r] = ids[best];
            out_score[qi * topk + r] = sc[best];
            ids[best] = -1;
        }
    }
}

#if defined(__CUDACC__)
__global__ void __launch_bounds__(256)
hybrid_rescore_kernel(int64_t total, int64_t pool, const double *__restrict__ q, int qpitch, const double *__restrict__ items,
                      int pitch, int f, int64_t row0, const double *__restrict__ norm_x, const double *__restrict__ lam_x,
                      const double *__restrict__ norm_q, const double *__restrict__ lam_q, double tau,
                      const int64_t *__restrict__ pool_idx, double *__restrict__ pool_score)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
        rescore_slot(t, pool, q, qpitch, items, pitch, f, row0, norm_x, lam_x, norm_q, lam_q, tau, pool_idx, pool_score);
}

__global__ void __launch_bounds__(128)
hybrid_select_kernel(int64_t nq, int64_t pool, int64_t topk, int64_t *pool_idx, const double *__restrict__ pool_score,
                     int64_t *__restrict__ out_idx, double *__restrict__ out_score)
{
    for (int64_t qi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; qi < nq; qi += (int64_t)gridDim.x * blockDim.x)
        select_query(qi, pool, topk, pool_idx, pool_score, out_idx, out_score);
}
#endif

}  // namespace asp_hybrid
