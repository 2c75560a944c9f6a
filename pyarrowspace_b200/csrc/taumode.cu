// taumode.cu -- K3: per-vector taumode lambda (items at build time, queries at search time).
//   E = x^T L x / x^T x,  tau = max(median(x), 1e-9),  lambda = E / (E + tau)
// (TAUMODE.md:18-19,24-25; replaces the lambda pass of
// ArrowSpaceBuilder::build and ArrowSpace::prepare_query_item, /root/reference/src/lib.rs:289,154).
//
// Formulation: L = D - W is symmetric, so  x^T L x = sum_a x_a (deg_a x_a - 2 sum_{b>a} w_ab x_b):
// only the strictly-upper adjacency is walked (half the gathers of a CSR SpMM).
// Bound: HBM (8*n*f bytes, X read once) while nnz(L)/f is small; for k = 25 the gathers from the
// shared-memory X tile dominate (8 bytes per nonzero per item) -- see DESIGN.md.
//
// Two kernels:
//  median_kernel   one warp per vector, values in registers, warp-cooperative quickselect (counts through
//                  __reduce_add_sync); streaming, high occupancy
//  taumode_kernel  CTA = 256 threads, one tile of T = 16*R items at a time, grid-stride (persistent):
//   A. warp w loads item rows (coalesced) and stores them transposed into shared memory
//      xs[feature][item] (row stride T+1: conflict free both ways)
//   A' thread t < T: left-to-right sum of squares (the norm the search kernel divides by; same
//      order as the oracle) and, for tau_mode = mean, the left-to-right sum
//   B. thread (part p = tid/16, lane-group g = tid%16) owns items g+16r (r < R) and the graph rows
//      a = p, p+16, ...; the upper adjacency is staged through shared memory in chunks; each
//      nonzero costs one broadcast LDS (col, weight) and R conflict-free LDS.64 of x
//   C. the 16 partial energies of an item are summed in part order; E, tau, lambda written.
#include "common.cuh"

#include <math.h>
#include <stdlib.h>

namespace {

constexpr int CH_NNZ_MAX = 1536; // upper-adjacency entries staged per chunk (>= longest row); per-variant value CHN below
constexpr int CH_ROWS = 256;
constexpr double TAU_FLOOR = 1e-9;

struct TmChunk { int row_begin; int row_end; };

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

// rank-th smallest (0-based) of the n finite values spread over the warp's registers
// (v[j] of lane l is element l + 32 j; absent elements are +inf).  Also returns the number of
// elements <= result in *count_le.
template <int FPL>
__device__ double warp_select(const double (&v)[FPL], int rank, int lane, uint32_t seed, int *count_le)
{
    double lo = -INFINITY, hi = INFINITY;
    int nbelow = 0;                       // elements <= lo
    int ca = 0;                           // this lane's elements inside (lo, hi)
#pragma unroll
    for (int j = 0; j < FPL; ++j) ca += (v[j] < hi) ? 1 : 0;
    for (int iter = 0;; ++iter) {
        // inclusive scan of the per-lane active counts
        int incl = ca;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total <= 0) { *count_le = nbelow; return NAN; }       // only with NaN / inf data
        const int pick = (int)(hash32(seed + iter) % (uint32_t)total);
        const unsigned ballot = __ballot_sync(0xffffffffu, incl > pick);
        const int owner = __ffs(ballot) - 1;
        double pv = 0.0;
        if (lane == owner) {
            const int want = pick - (incl - ca);
            int seen = 0;
#pragma unroll
            for (int j = 0; j < FPL; ++j) {
                const bool act = (v[j] > lo) && (v[j] < hi);
                if (act && seen == want) pv = v[j];
                seen += act ? 1 : 0;
            }
        }
        pv = __shfl_sync(0xffffffffu, pv, owner);
        int clt = 0, ceq = 0;
#pragma unroll
        for (int j = 0; j < FPL; ++j) {
            clt += ((v[j] > lo) && (v[j] < pv)) ? 1 : 0;
            ceq += (v[j] == pv) ? 1 : 0;
        }
        const int tlt = __reduce_add_sync(0xffffffffu, clt);
        const int teq = __reduce_add_sync(0xffffffffu, ceq);
        if (rank < nbelow + tlt) {
            hi = pv;
            ca = clt;
        } else if (rank < nbelow + tlt + teq) {
            *count_le = nbelow + tlt + teq;
            return pv;
        } else {
            lo = pv;
            nbelow += tlt + teq;
            ca = ca - clt - ceq;
        }
    }
}

template <int FPL>
__device__ double warp_median(const double (&v)[FPL], int n, int lane, uint32_t seed)
{
    int cle = 0;
    if (n & 1) return warp_select<FPL>(v, n / 2, lane, seed, &cle);
    const double vlo = warp_select<FPL>(v, n / 2 - 1, lane, seed, &cle);
    double vhi = vlo;
    if (cle < n / 2 + 1) {               // the next order statistic is the smallest element > vlo
        double m = INFINITY;
#pragma unroll
        for (int j = 0; j < FPL; ++j)
            if (v[j] > vlo && v[j] < m) m = v[j];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, off));
        vhi = m;
    }
    return 0.5 * (vlo + vhi);            // oracle.c median_of: 0.5 * (s[n/2-1] + s[n/2])
}

// K3a: per-vector median (tau before flooring), one warp per vector, grid-stride.  A pure streaming pass with
// small footprint (no shared memory, FPL f64 registers per lane): many resident warps hide the shuffle latency
// of the selection, which the tile kernel below (1 CTA / SM, 222 KB of shared memory) cannot.
template <int FPL>
__global__ void __launch_bounds__(256)
median_kernel(const double *__restrict__ x, int64_t n, int f, int pitch, int use_abs, double *__restrict__ out_median)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t item = warp; item < n; item += nwarps) {
        const double *row = x + item * pitch;
        double v[FPL];
#pragma unroll
        for (int j = 0; j < FPL; ++j) {
            const int ff = lane + 32 * j;
            v[j] = (ff < f) ? row[ff] : INFINITY;
            if (use_abs && ff < f) v[j] = fabs(v[j]);
        }
        const double med = warp_median<FPL>(v, f, lane, (uint32_t)(item * 2654435761ULL));
        if (lane == 0) out_median[item] = med;
    }
}

template <int FPL, int R, int CHN, int THR>
__global__ void __launch_bounds__(THR, 1)
taumode_kernel(const double *__restrict__ x, int64_t n, int f, int pitch, const int32_t *__restrict__ uptr,
               const int32_t *__restrict__ ucol, const double *__restrict__ uval, const double *__restrict__ deg,
               const TmChunk *__restrict__ chunks, int nchunks, int tau_mode, double tau_fixed,
               const double *__restrict__ medians, double *__restrict__ out_energy, double *__restrict__ out_tau, double *__restrict__ out_lambda,
               double *__restrict__ out_norm, double *__restrict__ out_inv_norm, int *zero_flag)
{
    constexpr int T = 16 * R;
    constexpr int XS = T + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *xs = reinterpret_cast<double *>(smem_raw);                 // f * XS
    double *s_val = xs + (size_t)f * XS;                               // CHN   (aliased by red[(THR / 16)][T])
    int32_t *s_col = reinterpret_cast<int32_t *>(s_val + CHN);         // CHN
    int32_t *s_rptr = s_col + CHN;                                     // CH_ROWS + 1
    double *s_deg = reinterpret_cast<double *>(s_rptr + CH_ROWS + 2);  // CH_ROWS
    double *s_tau = s_deg + CH_ROWS;                                   // T
    double *s_n2 = s_tau + T;                                          // T
    double *red = s_val;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = threadIdx.x & 15, p = threadIdx.x >> 4;
    const int64_t ntiles = (n + T - 1) / T;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t item0 = tile * T;
        __syncthreads();                                               // previous tile fully consumed

        // ---- A: load, transpose into shared memory
        for (int t = warp; t < T; t += THR / 32) {
            const int64_t item = item0 + t;
            double v[FPL];
            if (item < n) {
                const double *row = x + item * pitch;
#pragma unroll
                for (int j = 0; j < FPL; ++j) {
                    const int ff = lane + 32 * j;
                    v[j] = (ff < f) ? row[ff] : INFINITY;
                }
            } else {
#pragma unroll
                for (int j = 0; j < FPL; ++j) v[j] = (lane + 32 * j < f) ? 0.0 : INFINITY;
            }
#pragma unroll
            for (int j = 0; j < FPL; ++j) {
                const int ff = lane + 32 * j;
                if (ff < f) xs[ff * XS + t] = v[j];
            }
        }
        __syncthreads();

        // ---- A': left-to-right sums (norm^2; mean)
        if (threadIdx.x < T) {
            const int t = threadIdx.x;
            double n2 = 0.0, sm = 0.0;
#pragma unroll 8
            for (int ff = 0; ff < f; ++ff) {                           // loads and products run ahead; only the adds are a chain
                const double xv = xs[ff * XS + t];
                n2 = __dadd_rn(n2, __dmul_rn(xv, xv));
                sm = __dadd_rn(sm, xv);
            }
            s_n2[t] = n2;
            double tau;
            if (tau_mode == ASP_TAU_MEAN) tau = sm / (double)f;
            else if (tau_mode == ASP_TAU_FIXED) tau = tau_fixed;
            else tau = (item0 + t < n) ? medians[item0 + t] : 1.0;
            s_tau[t] = (tau > TAU_FLOOR) ? tau : TAU_FLOOR;
        }

        // ---- B: x^T L x through the strictly-upper adjacency
        double en[R];
#pragma unroll
        for (int r = 0; r < R; ++r) en[r] = 0.0;
        for (int c = 0; c < nchunks; ++c) {
            const int ra0 = chunks[c].row_begin, ra1 = chunks[c].row_end;
            const int e0 = uptr[ra0], e1 = uptr[ra1];
            __syncthreads();
            for (int i = threadIdx.x; i < e1 - e0; i += THR) { s_col[i] = ucol[e0 + i]; s_val[i] = uval[e0 + i]; }
            for (int i = threadIdx.x; i <= ra1 - ra0; i += THR) s_rptr[i] = uptr[ra0 + i] - e0;
            for (int i = threadIdx.x; i < ra1 - ra0; i += THR) s_deg[i] = deg[ra0 + i];
            __syncthreads();
            for (int a = ra0 + p; a < ra1; a += (THR / 16)) {
                double xa[R], s[R];
#pragma unroll
                for (int r = 0; r < R; ++r) { xa[r] = xs[a * XS + g + 16 * r]; s[r] = 0.0; }
                const int jb = s_rptr[a - ra0], je = s_rptr[a - ra0 + 1];
                for (int j = jb; j < je; ++j) {
                    const int b = s_col[j];
                    const double w = s_val[j];
#pragma unroll
                    for (int r = 0; r < R; ++r) s[r] = fma(w, xs[b * XS + g + 16 * r], s[r]);
                }
                const double dg = s_deg[a - ra0];
#pragma unroll
                for (int r = 0; r < R; ++r) en[r] = fma(xa[r], fma(dg, xa[r], -2.0 * s[r]), en[r]);
            }
        }
        __syncthreads();

        // ---- C: reduce the parts in order, finish
#pragma unroll
        for (int r = 0; r < R; ++r) red[p * T + g + 16 * r] = en[r];
        __syncthreads();
        if (threadIdx.x < T) {
            const int t = threadIdx.x;
            const int64_t item = item0 + t;
            if (item < n) {
                double num = 0.0;
                for (int q = 0; q < (THR / 16); ++q) num += red[q * T + t];
                const double n2 = s_n2[t];
                const double tau = s_tau[t];
                double e = NAN, lam = NAN;
                if (n2 == 0.0) atomicExch(zero_flag, 1);                 // TAUMODE.md:13
                else { e = num / n2; lam = e / (e + tau); }
                if (out_energy) out_energy[item] = e;
                if (out_tau) out_tau[item] = tau;
                if (out_lambda) out_lambda[item] = lam;
                const double nr = sqrt(n2);
                if (out_norm) out_norm[item] = nr;
                if (out_inv_norm) out_inv_norm[item] = (nr > 0.0) ? 1.0 / nr : 0.0;
            }
        }
    }
}

template <int FPL, int R, int CHN, int THR>
int launch_tm(asp_ctx *ctx, const asp_graph *g, const asp_switches *sw, const double *x, int64_t n, int f, int pitch,
              const TmChunk *d_chunks, int nchunks, double *oe, double *ot, double *ol, double *on, double *oi,
              int *zero_flag)
{
    constexpr int T = 16 * R;
    double *medians = nullptr;
    if (sw->tau_mode == ASP_TAU_MEDIAN || sw->tau_mode == ASP_TAU_MEDIAN_ABS) {
        ASP_CUDA(cudaMallocAsync(&medians, sizeof(double) * n, ctx->stream));
        const int64_t want = asp_ceil_div(n, 8);
        const int mgrid = (int)(want < (int64_t)ctx->num_sms * 8 ? want : (int64_t)ctx->num_sms * 8);
        median_kernel<FPL><<<mgrid, 256, 0, ctx->stream>>>(x, n, f, pitch, sw->tau_mode == ASP_TAU_MEDIAN_ABS ? 1 : 0, medians);
        ASP_CUDA(cudaGetLastError());
        ASP_LAUNCHED(ctx);
    }
    static_assert((size_t)CHN * 12 >= (size_t)(THR / 16) * T * 8, "the part reduction aliases the staged weights and columns");
    const size_t smem = (size_t)f * (T + 1) * 8 + (size_t)CHN * 12 + (CH_ROWS + 2) * 4 + (size_t)CH_ROWS * 8 +
                        (size_t)T * 16 + 64;
    if (smem > 227 * 1024) ASP_FAIL(ASP_ERR_UNSUPPORTED, "taumode kernel: %d features do not fit in shared memory", f);
    auto kern = taumode_kernel<FPL, R, CHN, THR>;
    ASP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;                                                       // resident CTAs per SM: their load / gather phases overlap
    ASP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THR, smem));
    if (occ < 1) occ = 1;
    const int64_t ntiles = (n + T - 1) / T;
    const int64_t slots = (int64_t)ctx->num_sms * occ;
    const int grid = (int)(ntiles < slots ? ntiles : slots);
    kern<<<grid, THR, smem, ctx->stream>>>(x, n, f, pitch, g->d_uptr, g->d_ucol, g->d_uval, g->d_deg, d_chunks,
                                                  nchunks, sw->tau_mode, sw->tau_fixed, medians, oe, ot, ol, on, oi, zero_flag);
    ASP_CUDA(cudaGetLastError());
    ASP_LAUNCHED(ctx);
    if (medians) ASP_CUDA(cudaFreeAsync(medians, ctx->stream));
    return ASP_OK;
}

}  // namespace

int asp_launch_taumode(asp_ctx *ctx, const asp_graph *g, const asp_switches *sw, const double *x_dev, int64_t n,
                       int32_t f, int32_t pitch, double *out_energy, double *out_tau, double *out_lambda,
                       double *out_norm, double *out_inv_norm, int *zero_flag_dev)
{
    if (n == 0) return ASP_OK;
    if (f != g->nnodes) ASP_FAIL(ASP_ERR_ARG, "vector length %d must equal the graph's node count %lld", f, (long long)g->nnodes);
    if (!g->d_uptr) ASP_FAIL(ASP_ERR_ARG, "graph has no upper adjacency (not a feature graph)");
    // chunk the upper adjacency by rows: <= chn entries and <= CH_ROWS rows per chunk.  The table depends on the graph
    // only: it is built and uploaded once (a per-call H2D copy would queue behind the query uploads of a pipelined search)
    if (!g->d_tm_chunks) {
        const int chn = CH_NNZ_MAX;
        std::vector<TmChunk> chunks;
        std::vector<int32_t> uptr(g->nnodes + 1);
        // host mirror of uptr: rebuild from the host CSR (cheap, f rows)
        int32_t acc = 0;
        uptr[0] = 0;
        for (int64_t a = 0; a < g->nnodes; ++a) {
            for (int64_t j = g->h_indptr[a]; j < g->h_indptr[a + 1]; ++j)
                if (g->h_indices[j] > a) ++acc;
            uptr[a + 1] = acc;
        }
        int a0 = 0;
        while (a0 < g->nnodes) {
            int a1 = a0;
            while (a1 < g->nnodes && (a1 - a0) < CH_ROWS && (uptr[a1 + 1] - uptr[a0]) <= chn) ++a1;
            if (a1 == a0) ASP_FAIL(ASP_ERR_UNSUPPORTED, "graph row %d has more than %d upper neighbours", a0, chn);
            chunks.push_back(TmChunk{a0, a1});
            a0 = a1;
        }
        void *d = nullptr;
        ASP_CUDA(cudaMallocAsync(&d, sizeof(TmChunk) * chunks.size(), ctx->stream));
        ASP_CUDA(cudaMemcpyAsync(d, chunks.data(), sizeof(TmChunk) * chunks.size(), cudaMemcpyHostToDevice, ctx->stream));
        ASP_CUDA(cudaStreamSynchronize(ctx->stream));     // chunks vector goes out of scope
        g->d_tm_chunks = d;
        g->n_tm_chunks = (int)chunks.size();
    }
    const TmChunk *d_chunks = static_cast<const TmChunk *>(g->d_tm_chunks);
    int rc;
    const int nch = g->n_tm_chunks;
    if (f <= 128)       rc = launch_tm<4, 4, CH_NNZ_MAX, 256>(ctx, g, sw, x_dev, n, f, pitch, d_chunks, nch, out_energy, out_tau, out_lambda, out_norm, out_inv_norm, zero_flag_dev);
    else if (f <= 384)  rc = launch_tm<12, 4, CH_NNZ_MAX, 512>(ctx, g, sw, x_dev, n, f, pitch, d_chunks, nch, out_energy, out_tau, out_lambda, out_norm, out_inv_norm, zero_flag_dev);
    else if (f <= 768)  rc = launch_tm<24, 2, CH_NNZ_MAX, 256>(ctx, g, sw, x_dev, n, f, pitch, d_chunks, nch, out_energy, out_tau, out_lambda, out_norm, out_inv_norm, zero_flag_dev);
    else if (f <= 1500) rc = launch_tm<48, 1, CH_NNZ_MAX, 256>(ctx, g, sw, x_dev, n, f, pitch, d_chunks, nch, out_energy, out_tau, out_lambda, out_norm, out_inv_norm, zero_flag_dev);
    else { rc = ASP_ERR_UNSUPPORTED; asp_set_error("taumode kernel supports at most 1500 features (got %d)", f); }
    return rc;
}
