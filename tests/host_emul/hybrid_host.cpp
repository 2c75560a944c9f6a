// CPU walk of the thread grid of pyarrowspace_b200/csrc/hybrid.cuh (test infrastructure: tests/test_hybrid_host.py).
// The kernels' per-thread bodies are plain functions; this file calls them for every thread index a launch would cover, in
// an order that differs from the GPU's only where threads are independent.  Build: g++ -O2 -ffp-contract=off -shared -fPIC.
#include "../../pyarrowspace_b200/csrc/hybrid.cuh"

extern "C" void hyb_emulate(int64_t nq, int64_t pool, int64_t topk, const double *q, int qpitch, const double *items, int pitch,
                            int f, int64_t row0, const double *norm_x, const double *lam_x, const double *norm_q,
                            const double *lam_q, double tau, int64_t *pool_idx, double *pool_score, int64_t *out_idx,
                            double *out_score)
{
    const int64_t total = nq * pool;
    // hybrid_rescore_kernel: grid-stride loop, 256 threads per block, a grid smaller than the work
    const int64_t threads_r = 3 * 256;
    for (int64_t tid = threads_r - 1; tid >= 0; --tid)                       // any thread order
        for (int64_t t = tid; t < total; t += threads_r)
            asp_hybrid::rescore_slot(t, pool, q, qpitch, items, pitch, f, row0, norm_x, lam_x, norm_q, lam_q, tau, pool_idx,
                                     pool_score);
    // hybrid_select_kernel
    const int64_t threads_s = 2 * 128;
    for (int64_t tid = threads_s - 1; tid >= 0; --tid)
        for (int64_t qi = tid; qi < nq; qi += threads_s)
            asp_hybrid::select_query(qi, pool, topk, pool_idx, pool_score, out_idx, out_score);
}
