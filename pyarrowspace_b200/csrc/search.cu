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
#include "common.cuh"
#include "ptx.cuh"
#include "warp_sort.cuh"

#include <math.h>

namespace {

using asp::Cand;

constexpr int QT = 128;            // queries per CTA tile
constexpr int IT = 128;            // items per tile
constexpr int KSTEP = 16;          // features per pipeline stage (4 DMMA k-slabs)
constexpr int OPER_DOUBLES = (KSTEP / 4) * 128 * 4;       // 2048 doubles = 16 KB
constexpr int STAGE_DOUBLES_S = 2 * OPER_DOUBLES;
constexpr int MMA_WARPS = 8;

// ============================================================================ stage 1: GEMM

template <int LIST>
struct ListSmem {
    static constexpr int CAP = 2 * LIST;
    double sc[QT * CAP];
    int32_t ix[QT * CAP];
    int32_t cnt[QT];
    double theta[QT];
};

// Sort the CAP-slot buffer of `row`, keep the best LIST, refresh the threshold.  Whole warp.
template <int LIST>
__device__ __noinline__ void compact_row(ListSmem<LIST> *ls, int row, int lane)
{
    constexpr int CAP = 2 * LIST;
    constexpr int NPL = CAP / 32;
    const int cnt = ls->cnt[row];
    Cand e[NPL];
#pragma unroll
    for (int t = 0; t < NPL; ++t) {
        const int i = lane + 32 * t;
        if (i < cnt) { e[t].s = ls->sc[row * CAP + i]; e[t].i = ls->ix[row * CAP + i]; }
        else e[t] = asp::cand_empty();
    }
    asp::warp_sort_best_first<NPL>(e, lane);
#pragma unroll
    for (int t = 0; t < NPL; ++t) {
        const int i = lane + 32 * t;
        if (i < LIST) { ls->sc[row * CAP + i] = e[t].s; ls->ix[row * CAP + i] = e[t].i; }
    }
    // element LIST-1 lives in lane (LIST-1)%32, slot (LIST-1)/32
    const double last = __shfl_sync(0xffffffffu, e[(LIST - 1) / 32].s, (LIST - 1) & 31);
    if (lane == 0) {
        ls->cnt[row] = cnt < LIST ? cnt : LIST;
        ls->theta[row] = (cnt >= LIST) ? last : -INFINITY;
    }
    __syncwarp();
}

template <int LIST, int STAGES, bool USE_TMA>
__global__ void __launch_bounds__(MMA_WARPS * 32, 1)
search_gemm_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x,
                   const double *__restrict__ q, const double *__restrict__ items, int64_t nq, int64_t n_local, int fp,
                   const double *__restrict__ inv_nx, const double *__restrict__ lam_x,
                   const double *__restrict__ inv_nq, const double *__restrict__ lam_q, double tau,
                   int nchunks, double *__restrict__ cand_score, int32_t *__restrict__ cand_idx)
{
    constexpr int CAP = 2 * LIST;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *stages = reinterpret_cast<double *>(smem_raw);
    ListSmem<LIST> *ls = reinterpret_cast<ListSmem<LIST> *>(stages + STAGES * STAGE_DOUBLES_S);
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qb = blockIdx.x, chunk = blockIdx.y;
    const int64_t tiles_total = (n_local + IT - 1) / IT;
    const int64_t tile0 = (tiles_total * chunk) / nchunks;               // even split of the item tiles
    const int64_t ntiles = (tiles_total * (chunk + 1)) / nchunks - tile0;
    const int ksteps = (fp + KSTEP - 1) / KSTEP;
    const int64_t total_it = ntiles * ksteps;

    for (int i = threadIdx.x; i < QT; i += blockDim.x) { ls->cnt[i] = 0; ls->theta[i] = -INFINITY; }
    if (USE_TMA && threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            asp::mbar_init(&full_bar[s], 1);
            asp::mbar_init(&empty_bar[s], MMA_WARPS);
        }
        asp::fence_barrier_init();
    }
    __syncthreads();

    // TMA producer = thread 0, inline: it runs STAGES-1 iterations ahead of the DMMA loop.
    auto tma_issue = [&](int64_t it) {
        const int s = (int)(it % STAGES);
        if (it >= STAGES) asp::mbar_wait(&empty_bar[s], (uint32_t)(((it / STAGES) - 1) & 1));
        double *dst = stages + s * STAGE_DOUBLES_S;
        const int64_t jt = it / ksteps;
        const int kk = (int)(it % ksteps);
        asp::mbar_arrive_expect_tx(&full_bar[s], STAGE_DOUBLES_S * 8u);
        asp::tma_load_3d(dst, &tmap_q, &full_bar[s], 0, qb * QT, kk * (KSTEP / 4));
        asp::tma_load_3d(dst + OPER_DOUBLES, &tmap_x, &full_bar[s], 0, (int)((tile0 + jt) * IT), kk * (KSTEP / 4));
    };
    if (USE_TMA && threadIdx.x == 0) {
        asp::tma_prefetch_desc(&tmap_q);
        asp::tma_prefetch_desc(&tmap_x);
        for (int64_t it = 0; it < STAGES - 1 && it < total_it; ++it) tma_issue(it);
    }

    // ===== consumers: warp w owns query rows [16w, 16w+16) x all 128 item columns of the tile
    const int rowA = warp * 16 + (lane >> 2);       // + 8*mt
    const int a_off = (rowA * 4) + (lane & 3);      // + ks*512 + mt*32
    const int b_off = ((lane >> 2) * 4) + (lane & 3);   // + ks*512 + nt*32

    // per-row constants of this lane's two rows
    double rq[2], lq[2], theta[2];
    bool qvalid[2];
    int rows[2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        rows[mt] = rowA + 8 * mt;
        const int64_t gq = (int64_t)qb * QT + rows[mt];
        qvalid[mt] = gq < nq;
        rq[mt] = qvalid[mt] ? tau * inv_nq[gq] : 0.0;
        lq[mt] = qvalid[mt] ? lam_q[gq] : 0.0;
        theta[mt] = -INFINITY;
    }
    const double beta = 1.0 - tau;

    auto load_stage_cp_async = [&](int64_t it) {
        const int s = (int)(it % STAGES);
        double *dst = stages + s * STAGE_DOUBLES_S;
        const int64_t jt = it / ksteps;
        const int kk = (int)(it % ksteps);
        for (int op = 0; op < 2; ++op) {
            const double *base = op == 0 ? q : items;
            const int64_t row_base = op == 0 ? (int64_t)qb * QT : (tile0 + jt) * IT;
            const int64_t row_lim = op == 0 ? nq : n_local;
            for (int c = threadIdx.x; c < 128 * KSTEP / 2; c += MMA_WARPS * 32) {
                const int r = c / (KSTEP / 2);
                const int fl = (c % (KSTEP / 2)) * 2;         // feature inside the k-step
                const int fg = kk * KSTEP + fl;
                const bool valid = (row_base + r < row_lim) && (fg < fp);
                const double *src = valid ? base + (row_base + r) * fp + fg : base;
                asp::cp_async16(dst + op * OPER_DOUBLES + (((fl >> 2) * 128 + r) * 4 + (fl & 3)), src, valid);
            }
        }
    };
    if (!USE_TMA) {
        for (int64_t it = 0; it < STAGES - 1; ++it) {
            if (it < total_it) load_stage_cp_async(it);
            asp::cp_async_commit();
        }
    }

    for (int64_t jt = 0; jt < ntiles; ++jt) {
        double acc[2][16][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 16; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

        for (int kk = 0; kk < ksteps; ++kk) {
            const int64_t it = jt * ksteps + kk;
            const int s = (int)(it % STAGES);
            if (USE_TMA) {
                if (threadIdx.x == 0 && it + STAGES - 1 < total_it) tma_issue(it + STAGES - 1);
                asp::mbar_wait(&full_bar[s], (uint32_t)((it / STAGES) & 1));
            } else {
                asp::cp_async_wait<STAGES - 2>();
                __syncthreads();
                if (it + STAGES - 1 < total_it) load_stage_cp_async(it + STAGES - 1);
                asp::cp_async_commit();
            }
            const double *A = stages + s * STAGE_DOUBLES_S;
            const double *B = A + OPER_DOUBLES;
#pragma unroll
            for (int ks = 0; ks < KSTEP / 4; ++ks) {
                double a[2];
                a[0] = A[a_off + ks * 512];
                a[1] = A[a_off + ks * 512 + 32];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    double b[8];
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) b[nt] = B[b_off + ks * 512 + (half * 8 + nt) * 32];
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) {
                        asp::dmma884(acc[0][half * 8 + nt][0], acc[0][half * 8 + nt][1], a[0], b[nt]);
                        asp::dmma884(acc[1][half * 8 + nt][0], acc[1][half * 8 + nt][1], a[1], b[nt]);
                    }
                }
            }
            if (USE_TMA) {
                __syncwarp();
                if (lane == 0) asp::mbar_arrive(&empty_bar[s]);
            }
        }

        // ----- fused epilogue: score, threshold, push, compact (warp local)
        const int64_t item_base = (tile0 + jt) * IT;
#pragma unroll
        for (int nt = 0; nt < 16; ++nt) {
            const int64_t c0 = item_base + nt * 8 + 2 * (lane & 3);
            double inx[2], lmx[2];
            bool cvalid[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                cvalid[e] = (c0 + e) < n_local;
                inx[e] = cvalid[e] ? inv_nx[c0 + e] : 0.0;
                lmx[e] = cvalid[e] ? lam_x[c0 + e] : 0.0;
            }
            bool pushed = false;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double dot = acc[mt][nt][e];
                    const double cs = rq[mt] * dot * inx[e];
                    if (cvalid[e] && qvalid[mt] && (cs + beta > theta[mt])) {
                        const double sc = cs + beta / (1.0 + fabs(lq[mt] - lmx[e]));
                        if (sc > theta[mt]) {
                            const int slot = atomicAdd(&ls->cnt[rows[mt]], 1);
                            ls->sc[rows[mt] * CAP + slot] = sc;
                            ls->ix[rows[mt] * CAP + slot] = (int32_t)(c0 + e);
                            pushed = true;
                        }
                    }
                }
            }
            if (__any_sync(0xffffffffu, pushed)) {
                __syncwarp();
                const int myrow = warp * 16 + (lane & 15);
                unsigned need = __ballot_sync(0xffffffffu, (lane < 16) && (ls->cnt[myrow] > CAP - 8));
                while (need) {
                    const int rr = __ffs(need) - 1;
                    need &= need - 1;
                    compact_row<LIST>(ls, warp * 16 + rr, lane);
                }
                __syncwarp();
                theta[0] = ls->theta[rows[0]];
                theta[1] = ls->theta[rows[1]];
            }
        }
    }

    // ----- flush: final compaction of the warp's 16 rows, best LIST out
    __syncwarp();
    for (int rr = 0; rr < 16; ++rr) {
        const int row = warp * 16 + rr;
        compact_row<LIST>(ls, row, lane);
        const int64_t gq = (int64_t)qb * QT + row;
        if (gq < nq) {
            const int cnt = ls->cnt[row];
            for (int i = lane; i < LIST; i += 32) {
                const size_t o = ((size_t)gq * nchunks + chunk) * LIST + i;
                cand_score[o] = (i < cnt) ? ls->sc[row * CAP + i] : -INFINITY;
                cand_idx[o] = (i < cnt) ? ls->ix[row * CAP + i] : -1;
            }
        }
    }
}

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

__device__ __forceinline__ double seq_dot_row(const double *__restrict__ qv, const double *__restrict__ row, int f)
{
    double d = 0.0;
    int j = 0;
    for (; j + 4 <= f; j += 4) {
        const double2 x0 = *reinterpret_cast<const double2 *>(row + j);
        const double2 x1 = *reinterpret_cast<const double2 *>(row + j + 2);
        d = __dadd_rn(d, __dmul_rn(qv[j], x0.x));
        d = __dadd_rn(d, __dmul_rn(qv[j + 1], x0.y));
        d = __dadd_rn(d, __dmul_rn(qv[j + 2], x1.x));
        d = __dadd_rn(d, __dmul_rn(qv[j + 3], x1.y));
    }
    for (; j < f; ++j) d = __dadd_rn(d, __dmul_rn(qv[j], row[j]));
    return d;
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

__global__ void exact_scan_kernel(const double *__restrict__ q, int qpitch, const int32_t *__restrict__ slow_list, int slot,
                                  const double *__restrict__ items, int64_t n_local, int f, int pitch,
                                  const double *__restrict__ norm_x, const double *__restrict__ lam_x,
                                  const double *__restrict__ norm_q, const double *__restrict__ lam_q, double tau,
                                  double *__restrict__ scores)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *qs = reinterpret_cast<double *>(smem_raw);
    const int64_t qi = slow_list[slot];
    for (int j = threadIdx.x; j < f; j += blockDim.x) qs[j] = q[qi * qpitch + j];
    __syncthreads();
    const double nqv = norm_q[qi], lqv = lam_q[qi];
    for (int64_t it = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; it < n_local; it += (int64_t)gridDim.x * blockDim.x) {
        const double d = seq_dot_row(qs, items + it * pitch, f);
        scores[it] = exact_score(d, nqv, norm_x[it], tau, lqv, lam_x[it]);
    }
}

// single block: topk rounds of "best element strictly after the previous winner"
__global__ void exact_select_kernel(const double *__restrict__ scores, int64_t n_local, int64_t row0, int topk,
                                    const int32_t *__restrict__ slow_list, int slot, int64_t *__restrict__ out_idx,
                                    double *__restrict__ out_score)
{
    __shared__ double s_s[32];
    __shared__ int64_t s_i[32];
    __shared__ double prev_s;
    __shared__ int64_t prev_i;
    const int64_t qi = slow_list[slot];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { prev_s = INFINITY; prev_i = -1; }
    __syncthreads();
    for (int r = 0; r < topk; ++r) {
        const double ps = prev_s;
        const int64_t pi = prev_i;
        double bs = -INFINITY;
        int64_t bi = INT64_MAX;
        for (int64_t it = threadIdx.x; it < n_local; it += blockDim.x) {
            const double s = scores[it];
            const bool after_prev = (s < ps) || (s == ps && it > pi);
            if (after_prev && ((s > bs) || (s == bs && it < bi))) { bs = s; bi = it; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double os = __shfl_xor_sync(0xffffffffu, bs, off);
            const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if ((os > bs) || (os == bs && oi < bi)) { bs = os; bi = oi; }
        }
        if (lane == 0) { s_s[warp] = bs; s_i[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            bs = (lane < (int)(blockDim.x >> 5)) ? s_s[lane] : -INFINITY;
            bi = (lane < (int)(blockDim.x >> 5)) ? s_i[lane] : INT64_MAX;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double os = __shfl_xor_sync(0xffffffffu, bs, off);
                const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if ((os > bs) || (os == bs && oi < bi)) { bs = os; bi = oi; }
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
        auto k = search_gemm_kernel<LIST, STAGES, true>;
        ASP_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, MMA_WARPS * 32, smem, ctx->stream>>>(tmap_q, s->tmap_rows, q_dev, s->items, nq, s->n_local, s->fp,
                                                           s->inv_norms, s->lambdas, inv_nq, lam_q, tau, nchunks, cand_score,
                                                           cand_idx);
    } else {
        auto k = search_gemm_kernel<LIST, STAGES, false>;
        ASP_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, MMA_WARPS * 32, smem, ctx->stream>>>(tmap_q, s->tmap_rows, q_dev, s->items, nq, s->n_local, s->fp,
                                                     s->inv_norms, s->lambdas, inv_nq, lam_q, tau, nchunks, cand_score,
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
    reciprocal_kernel<<<(unsigned)(asp_ceil_div(nq, 256) < 1024 ? asp_ceil_div(nq, 256) : 1024), 256, 0, st>>>(qnorm_dev, nq, inv_nq);
    ASP_CUDA(cudaGetLastError());
    ASP_LAUNCHED(ctx);

    int32_t *slow_list = nullptr, *slow_count = nullptr;
    ASP_CUDA(cudaMallocAsync(&slow_list, sizeof(int32_t) * (nq + 1), st));
    ASP_CUDA(cudaMallocAsync(&slow_count, sizeof(int32_t), st));
    ASP_CUDA(cudaMemsetAsync(slow_count, 0, sizeof(int32_t), st));

    int nparts = 0;
    double *cand_score = nullptr;
    int32_t *cand_idx = nullptr;
    ASP_CUDA(cudaEventRecord(ctx->ev0, st));
    if (nq <= GV_MAXQ) {
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
    if (nslow > 0) {
        double *scores = nullptr;
        ASP_CUDA(cudaMallocAsync(&scores, sizeof(double) * s->n_local, st));
        for (int i = 0; i < nslow; ++i) {
            exact_scan_kernel<<<ctx->num_sms * 4, 256, (size_t)f * 8, st>>>(q_dev, qpitch, slow_list, i, s->items, s->n_local, f,
                                                                            s->fp, s->norms, s->lambdas, qnorm_dev,
                                                                            lambda_q_dev, tau, scores);
            ASP_CUDA(cudaGetLastError()); ASP_LAUNCHED(ctx);
            exact_select_kernel<<<1, 1024, 0, st>>>(scores, s->n_local, s->row0, (int)topk, slow_list, i, out_idx_dev,
                                                    out_score_dev);
            ASP_CUDA(cudaGetLastError()); ASP_LAUNCHED(ctx);
        }
        ASP_CUDA(cudaFreeAsync(scores, st));
    }
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
