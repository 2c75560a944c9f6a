// knn.cu -- K1, item orientation (nodes = items): eps-radius / k-NN lists on rectified-cosine distance
// for every item against every item (GRAPH_VARIABLES.md:7-8; SURVEY.md Appendix A2 `nodes = items`,
// A3-A4; the graph-build workload of BASELINE.json configs C4/C5).
//
// Bound: FP64 tensor pipe, 2*M^2*D FLOP (every ordered pair, as the reference's per-row scan does).
//
// Same two-stage contract as the search:
//   stage 1  search_gemm_kernel<MODE 1>: DMMA 128x128 tiles of X X^T fed by TMA, epilogue keeps per row the
//            LIST best rectified cosines >= 1 - eps - band (self excluded) in shared memory.
//   stage 2  knn_rescore_kernel: the candidates' distances are recomputed in the oracle's order
//            (left-to-right dot, d = 1 - max(0, dot/(|a||b|))), d <= eps is decided on those values, the k
//            smallest by (d, index) are kept.  Complete iff s~(LIST) < s~(k) - 2 band; otherwise the row takes
//            the exact scan (knn_exact_scan_kernel + knn_exact_select_kernel).
// Output: neighbour lists for csr.cu (K2).
#include "gemm_topk.cuh"

namespace {

using namespace asp_gemm;
using asp::Cand;

// left-to-right norms (the oracle's: sqrt of the sequential sum of squares), one thread per row
__global__ void row_norms_kernel(const double *__restrict__ items, int64_t n, int f, int pitch, double *__restrict__ norms,
                                 double *__restrict__ inv_norms)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double *row = items + i * pitch;
        const double n2 = seq_dot_row(row, row, f);
        const double nr = sqrt(n2);
        norms[i] = nr;
        inv_norms[i] = nr > 0.0 ? 1.0 / nr : 0.0;
    }
}

// the oracle's distance expression (oracle.c select_neighbours)
__device__ __forceinline__ double exact_dist(double dot, double na, double nb)
{
    double c = 0.0;
    if (na != 0.0 && nb != 0.0) c = __ddiv_rn(dot, __dmul_rn(na, nb));
    return __dsub_rn(1.0, c > 0.0 ? c : 0.0);
}

constexpr int KR_WARPS = 4;

template <int LIST>
__global__ void __launch_bounds__(KR_WARPS * 32)
knn_rescore_kernel(const double *__restrict__ items, int64_t n, int f, int pitch, const double *__restrict__ norms,
                   double eps, int kk, int nparts, const double *__restrict__ cand_score,
                   const int32_t *__restrict__ cand_idx, double delta, int32_t *__restrict__ out_idx,
                   double *__restrict__ out_dist, int32_t *__restrict__ out_cnt, int32_t *slow_list, int32_t *slow_count)
{
    constexpr int CAP = 2 * LIST;
    constexpr int NPL = CAP / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *xs_all = reinterpret_cast<double *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *xs = xs_all + (size_t)warp * f;
    const int64_t i = (int64_t)blockIdx.x * KR_WARPS + warp;
    if (i >= n) return;
    for (int j = lane; j < f; j += 32) xs[j] = items[i * pitch + j];
    __syncwarp();

    Cand best[NPL];
#pragma unroll
    for (int t = 0; t < NPL; ++t) best[t] = asp::cand_empty();
    for (int p = 0; p < nparts; ++p) {
        const size_t base = ((size_t)i * nparts + p) * LIST;
#pragma unroll
        for (int t = 0; t < NPL; ++t) {
            const int e = lane + 32 * t;
            if (e >= LIST) {
                const int32_t ci = cand_idx[base + e - LIST];
                if (ci >= 0) { best[t].s = cand_score[base + e - LIST]; best[t].i = ci; }
                else best[t] = asp::cand_empty();
            }
        }
        asp::warp_sort_best_first<NPL>(best, lane);
    }
    const double a_k = __shfl_sync(0xffffffffu, best[(kk - 1) / 32].s, (kk - 1) & 31);
    const double a_L = __shfl_sync(0xffffffffu, best[(LIST - 1) / 32].s, (LIST - 1) & 31);
    const bool complete = (a_L == -INFINITY) || (a_L < a_k - 2.0 * delta);
    if (!complete) {
        if (lane == 0) slow_list[atomicAdd(slow_count, 1)] = (int32_t)i;
        return;
    }
    const double ni = norms[i];
#pragma unroll
    for (int t = 0; t < NPL; ++t) {
        const int e = lane + 32 * t;
        if (e < LIST && best[t].i != 0x7fffffff) {
            const int64_t j = best[t].i;
            const double d = exact_dist(seq_dot_row(xs, items + j * pitch, f), ni, norms[j]);
            best[t].s = (d <= eps) ? -d : -INFINITY;            // GRAPH_VARIABLES.md:7
            if (!(d <= eps)) best[t].i = 0x7fffffff;
        } else best[t] = asp::cand_empty();
    }
    asp::warp_sort_best_first<NPL>(best, lane);                  // (-d desc, index asc) == (d asc, index asc)
    int nvalid = 0;
#pragma unroll
    for (int t = 0; t < NPL; ++t) nvalid += __popc(__ballot_sync(0xffffffffu, best[t].i != 0x7fffffff));
    const int keep = nvalid < kk ? nvalid : kk;
#pragma unroll
    for (int t = 0; t < NPL; ++t) {
        const int e = lane + 32 * t;
        if (e < keep) { out_idx[i * kk + e] = best[t].i; out_dist[i * kk + e] = -best[t].s; }
    }
    if (lane == 0) out_cnt[i] = keep;
}

// slow path: score[j] = -d(i,j) for valid neighbours, -inf otherwise
__global__ void knn_exact_scan_kernel(const double *__restrict__ items, int64_t n, int f, int pitch,
                                      const double *__restrict__ norms, double eps, const int32_t *__restrict__ slow_list,
                                      int slot, double *__restrict__ scores)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *xs = reinterpret_cast<double *>(smem_raw);
    const int64_t i = slow_list[slot];
    for (int j = threadIdx.x; j < f; j += blockDim.x) xs[j] = items[i * pitch + j];
    __syncthreads();
    const double ni = norms[i];
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const double d = exact_dist(seq_dot_row(xs, items + j * pitch, f), ni, norms[j]);
        scores[j] = (j != i && d <= eps) ? -d : -INFINITY;
    }
}

// single block: kk rounds of "best element strictly after the previous winner" in (score desc, index asc)
__global__ void knn_exact_select_kernel(const double *__restrict__ scores, int64_t n, int kk,
                                        const int32_t *__restrict__ slow_list, int slot, int32_t *__restrict__ out_idx,
                                        double *__restrict__ out_dist, int32_t *__restrict__ out_cnt)
{
    __shared__ double s_s[32];
    __shared__ int64_t s_i[32];
    __shared__ double prev_s;
    __shared__ int64_t prev_i;
    __shared__ int s_keep;
    const int64_t row = slow_list[slot];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { prev_s = INFINITY; prev_i = -1; s_keep = 0; }
    __syncthreads();
    for (int r = 0; r < kk; ++r) {
        const double ps = prev_s;
        const int64_t pi = prev_i;
        double bs = -INFINITY;
        int64_t bi = INT64_MAX;
        for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
            const double s = scores[j];
            if (s == -INFINITY) continue;
            const bool after_prev = (s < ps) || (s == ps && j > pi);
            if (after_prev && ((s > bs) || (s == bs && j < bi))) { bs = s; bi = j; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double os = __shfl_xor_sync(0xffffffffu, bs, off);
            const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (oi != INT64_MAX && ((os > bs) || (os == bs && oi < bi) || bi == INT64_MAX)) { bs = os; bi = oi; }
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
                if (oi != INT64_MAX && ((os > bs) || (os == bs && oi < bi) || bi == INT64_MAX)) { bs = os; bi = oi; }
            }
            if (lane == 0) {
                if (bi != INT64_MAX) {
                    out_idx[row * kk + r] = (int32_t)bi;
                    out_dist[row * kk + r] = -bs;
                    s_keep = r + 1;
                    prev_s = bs;
                    prev_i = bi;
                } else {
                    prev_s = -INFINITY;
                    prev_i = INT64_MAX;
                }
            }
        }
        __syncthreads();
        if (prev_i == INT64_MAX) break;
    }
    if (threadIdx.x == 0) out_cnt[row] = s_keep;
}

}  // namespace

int asp_item_knn(asp_space *s, const asp_graph_params *gp, asp_knn_lists *lists)
{
    constexpr int LIST = 32;
    constexpr int STAGES = 3;
    asp_ctx *ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    const int64_t n = s->n_local;
    const int f = s->f;
    if (n > 2147483647LL) ASP_FAIL(ASP_ERR_UNSUPPORTED, "item graph supports at most 2^31-1 items");
    int64_t kk = gp->k;
    if (kk > n - 1) kk = n - 1;
    if (kk < 0) kk = 0;
    if (kk > LIST - 4) ASP_FAIL(ASP_ERR_UNSUPPORTED, "item graph supports k <= %d (got %lld)", LIST - 4, (long long)gp->k);
    lists->m = n;
    lists->kk = (int32_t)(kk > 0 ? kk : 1);
    ASP_CUDA(cudaMallocAsync(&lists->idx, sizeof(int32_t) * (size_t)n * lists->kk, st));
    ASP_CUDA(cudaMallocAsync(&lists->dist, sizeof(double) * (size_t)n * lists->kk, st));
    ASP_CUDA(cudaMallocAsync(&lists->cnt, sizeof(int32_t) * (size_t)n, st));
    ASP_CUDA(cudaMemsetAsync(lists->cnt, 0, sizeof(int32_t) * (size_t)n, st));
    if (kk == 0) return ASP_OK;

    row_norms_kernel<<<(unsigned)(asp_ceil_div(n, 128) < 65535 ? asp_ceil_div(n, 128) : 65535), 128, 0, st>>>(
        s->items, n, f, s->fp, s->norms, s->inv_norms);
    ASP_CUDA(cudaGetLastError());
    ASP_LAUNCHED(ctx);

    // rounding band of one approximate cosine (two summation orders of f terms, both norms) + the 1 - c rounding
    const double u = 1.1102230246251565e-16;
    const double delta = (4.0 * f + 64.0) * u;
    const double smin = 1.0 - gp->eps - delta;

    const int64_t tiles_total = asp_ceil_div(n, IT);
    const int64_t qblocks = asp_ceil_div(n, QT);
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
    const int nparts = (int)best_chunks;

    double *cand_score = nullptr;
    int32_t *cand_idx = nullptr, *slow_list = nullptr, *slow_count = nullptr;
    ASP_CUDA(cudaMallocAsync(&cand_score, sizeof(double) * (size_t)n * nparts * LIST, st));
    ASP_CUDA(cudaMallocAsync(&cand_idx, sizeof(int32_t) * (size_t)n * nparts * LIST, st));
    ASP_CUDA(cudaMallocAsync(&slow_list, sizeof(int32_t) * (n + 1), st));
    ASP_CUDA(cudaMallocAsync(&slow_count, sizeof(int32_t), st));
    ASP_CUDA(cudaMemsetAsync(slow_count, 0, sizeof(int32_t), st));

    ASP_CUDA(cudaEventRecord(ctx->ev0, st));
    {
        const size_t smem = (size_t)STAGES * STAGE_DOUBLES_S * 8 + sizeof(ListSmem<LIST>) + 128;
        dim3 grid((unsigned)qblocks, nparts);
        if (ctx->use_tma) {
            auto k = search_gemm_kernel<LIST, STAGES, true, 1>;
            ASP_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k<<<grid, MMA_WARPS * 32, smem, st>>>(s->tmap_rows, s->tmap_rows, s->items, s->items, n, n, s->fp, s->inv_norms,
                                                   nullptr, s->inv_norms, nullptr, 1.0, smin, nparts, cand_score, cand_idx);
        } else {
            auto k = search_gemm_kernel<LIST, STAGES, false, 1>;
            ASP_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k<<<grid, MMA_WARPS * 32, smem, st>>>(s->tmap_rows, s->tmap_rows, s->items, s->items, n, n, s->fp, s->inv_norms,
                                                   nullptr, s->inv_norms, nullptr, 1.0, smin, nparts, cand_score, cand_idx);
        }
        ASP_CUDA(cudaGetLastError());
        ASP_LAUNCHED(ctx);
    }
    ASP_CUDA(cudaEventRecord(ctx->ev1, st));
    {
        const size_t smem = (size_t)KR_WARPS * f * 8;
        ASP_CUDA(cudaFuncSetAttribute(knn_rescore_kernel<LIST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn_rescore_kernel<LIST><<<(unsigned)asp_ceil_div(n, KR_WARPS), KR_WARPS * 32, smem, st>>>(
            s->items, n, f, s->fp, s->norms, gp->eps, (int)kk, nparts, cand_score, cand_idx, delta, lists->idx, lists->dist,
            lists->cnt, slow_list, slow_count);
        ASP_CUDA(cudaGetLastError());
        ASP_LAUNCHED(ctx);
    }
    int32_t nslow = 0;
    ASP_CUDA(cudaMemcpyAsync(&nslow, slow_count, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    ASP_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->stats["knn_stage1_ms"] = ms;
    ctx->stats["knn_slow_rows"] = nslow;
    if (nslow > 0) {
        double *scores = nullptr;
        ASP_CUDA(cudaMallocAsync(&scores, sizeof(double) * n, st));
        for (int i = 0; i < nslow; ++i) {
            knn_exact_scan_kernel<<<ctx->num_sms * 4, 256, (size_t)f * 8, st>>>(s->items, n, f, s->fp, s->norms, gp->eps,
                                                                                slow_list, i, scores);
            ASP_CUDA(cudaGetLastError()); ASP_LAUNCHED(ctx);
            knn_exact_select_kernel<<<1, 1024, 0, st>>>(scores, n, (int)kk, slow_list, i, lists->idx, lists->dist, lists->cnt);
            ASP_CUDA(cudaGetLastError()); ASP_LAUNCHED(ctx);
        }
        ASP_CUDA(cudaFreeAsync(scores, st));
    }
    ASP_CUDA(cudaFreeAsync(cand_score, st));
    ASP_CUDA(cudaFreeAsync(cand_idx, st));
    ASP_CUDA(cudaFreeAsync(slow_list, st));
    ASP_CUDA(cudaFreeAsync(slow_count, st));
    return ASP_OK;
}
