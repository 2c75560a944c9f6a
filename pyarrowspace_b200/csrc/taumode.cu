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
//  median_kernel   one warp per vector, order-preserving 64-bit keys in registers, MSB-first radix selection through a
//                  256-bin shared-memory histogram per warp (2-3 passes for F = 384); streaming, high occupancy
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

// ---- radix selection on order-preserving 64-bit keys (the streaming median kernel below)
// key(x) is monotone in x for every non-NaN double (-0.0 is folded into +0.0 first); absent elements carry the largest key.
__device__ __forceinline__ unsigned long long f64_key(double x)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(x + 0.0);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_f64(unsigned long long k)
{
    const unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// rank-th smallest key (0-based) of the n present keys held as k[j] of lane l = element l + 32 j.  MSB-first radix
// selection, 8 bits per pass through a 256-bin histogram in shared memory (one per warp); the bytes all present keys
// share are skipped (embeddings of one scale share sign, exponent and often the first mantissa bits), and the walk
// stops as soon as the selected bin holds one key -- two or three passes for F = 384.  *count_le = number of keys <= result.
template <int FPL>
__device__ unsigned long long warp_radix_select(const unsigned long long (&k)[FPL], int n, int rank, int lane,
                                                uint32_t *hist /* [256] of this warp */, int *count_le)
{
    // common leading bytes
    uint32_t dhi = 0, dlo = 0;
    const unsigned long long k0 = __shfl_sync(0xffffffffu, k[0], 0);            // element 0 is always present
#pragma unroll
    for (int j = 0; j < FPL; ++j) {
        const bool present = (lane + 32 * j) < n;
        const unsigned long long d = present ? (k[j] ^ k0) : 0ull;
        dhi |= (uint32_t)(d >> 32);
        dlo |= (uint32_t)d;
    }
    dhi = __reduce_or_sync(0xffffffffu, dhi);
    dlo = __reduce_or_sync(0xffffffffu, dlo);
    if ((dhi | dlo) == 0u) { *count_le = n; return k0; }                         // all keys equal
    const int lead = dhi ? __clz(dhi) : 32 + __clz(dlo);                         // identical leading bits
    // the first digit is the 8 most significant bits that VARY (not a byte boundary): the keys spread over all 256 bins,
    // so the shared-memory atomics of the first -- and only populous -- pass hardly collide
    const int top = 64 - lead;                                                   // low bits that may differ, >= 1
    int shift = top > 8 ? top - 8 : 0;
    int width = top - shift;
    unsigned long long prefix = (top == 64) ? 0ull : (k0 >> top) << top;
    unsigned long long mask = (top == 64) ? 0ull : ~0ull << top;
    int below = 0;                                                               // keys smaller than every key under the prefix
    int r = rank;
    for (;;) {
#pragma unroll
        for (int b = 0; b < 8; ++b) hist[lane + 32 * b] = 0u;
        __syncwarp();
#pragma unroll
        for (int j = 0; j < FPL; ++j) {
            const bool act = ((lane + 32 * j) < n) && ((k[j] & mask) == prefix);
            if (act) atomicAdd(&hist[(uint32_t)(k[j] >> shift) & ((1u << width) - 1u)], 1u);
        }
        __syncwarp();
        // lane l owns bins 8l .. 8l+7
        uint32_t c[8];
        const uint4 h0 = *reinterpret_cast<const uint4 *>(hist + 8 * lane), h1 = *reinterpret_cast<const uint4 *>(hist + 8 * lane + 4);
        c[0] = h0.x; c[1] = h0.y; c[2] = h0.z; c[3] = h0.w; c[4] = h1.x; c[5] = h1.y; c[6] = h1.z; c[7] = h1.w;
        uint32_t tot = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) tot += c[b];
        uint32_t incl = tot;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        const unsigned owners = __ballot_sync(0xffffffffu, incl > (uint32_t)r);
        const int owner = __ffs(owners) - 1;                                     // first lane whose inclusive count exceeds r
        uint32_t run = incl - tot, digit = 0, cnt = 0, before = 0;
        if (lane == owner) {
            bool found = false;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const bool here = !found && (run + c[b] > (uint32_t)r);
                if (here) { digit = 8 * lane + b; cnt = c[b]; before = run; found = true; }
                run += c[b];
            }
        }
        digit = __shfl_sync(0xffffffffu, digit, owner);
        cnt = __shfl_sync(0xffffffffu, cnt, owner);
        before = __shfl_sync(0xffffffffu, before, owner);
        below += (int)before;
        r -= (int)before;
        prefix |= (unsigned long long)digit << shift;
        mask |= (unsigned long long)((1u << width) - 1u) << shift;
        __syncwarp();                                                            // the histogram is reused by the next pass
        if (shift == 0) { *count_le = below + (int)cnt; return prefix; }         // cnt equal keys
        if (cnt == 1u) {                                                         // one key left under the prefix: fetch it
            unsigned long long mine = 0ull;
            bool have = false;
#pragma unroll
            for (int j = 0; j < FPL; ++j) {
                const bool act = ((lane + 32 * j) < n) && ((k[j] & mask) == prefix);
                if (act) { mine = k[j]; have = true; }
            }
            const int src = __ffs(__ballot_sync(0xffffffffu, have)) - 1;
            *count_le = below + 1;
            return __shfl_sync(0xffffffffu, mine, src);
        }
        const int next = shift > 8 ? shift - 8 : 0;
        width = shift - next;
        shift = next;
    }
}

template <int FPL>
__device__ double warp_median_radix(const unsigned long long (&k)[FPL], int n, int lane, uint32_t *hist)
{
    int cle = 0;
    if (n & 1) return key_f64(warp_radix_select<FPL>(k, n, n / 2, lane, hist, &cle));
    const unsigned long long klo = warp_radix_select<FPL>(k, n, n / 2 - 1, lane, hist, &cle);
    unsigned long long khi = klo;
    if (cle < n / 2 + 1) {                   // the next order statistic is the smallest key > klo
        unsigned long long m = ~0ull;
#pragma unroll
        for (int j = 0; j < FPL; ++j)
            if ((lane + 32 * j) < n && k[j] > klo && k[j] < m) m = k[j];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, m, off);
            m = (o < m) ? o : m;
        }
        khi = m;
    }
    return 0.5 * (key_f64(klo) + key_f64(khi));   // oracle.c median_of: 0.5 * (s[n/2-1] + s[n/2])
}

// K3a: per-vector median (tau before flooring), one warp per vector, grid-stride.  A pure streaming pass with a small
// footprint (1 KB of shared memory per warp, FPL 64-bit keys per lane): many resident warps hide the latency of the
// selection, which the tile kernel below (1 CTA / SM, 222 KB of shared memory) cannot.
// PF (ASP_MEDIAN_PREFETCH=1, not yet the default: written after the GPU budget of round 1 was spent): the row of the warp's
// NEXT item is requested before the selection on the current one, so the loads -- 59 % of this kernel's stall samples -- overlap
// the selection instead of preceding it.  Costs FPL more 64-bit registers (3 instead of 4 CTAs per SM at F <= 384).
template <int FPL, bool PF>
__global__ void __launch_bounds__(256, FPL <= 12 ? (PF ? 3 : 4) : 1)     // 64 registers: 32 warps per SM hide the selection's latency
median_kernel(const double *__restrict__ x, int64_t n, int f, int pitch, int use_abs, double *__restrict__ out_median)
{
    __shared__ __align__(16) uint32_t s_hist[8][256];
    const int lane = threadIdx.x & 31;
    uint32_t *hist = s_hist[threadIdx.x >> 5];
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double nv[FPL];                                                      // PF: the next item's values, in flight
    if (PF && warp < n) {
        const double *row = x + warp * pitch;
#pragma unroll
        for (int j = 0; j < FPL; ++j) { const int ff = lane + 32 * j; nv[j] = (ff < f) ? row[ff] : 0.0; }
    }
    for (int64_t item = warp; item < n; item += nwarps) {
        const double *row = x + item * pitch;
        unsigned long long k[FPL];
#pragma unroll
        for (int j = 0; j < FPL; ++j) {
            const int ff = lane + 32 * j;
            double v = PF ? nv[j] : ((ff < f) ? row[ff] : 0.0);
            if (use_abs) v = fabs(v);
            k[j] = (ff < f) ? f64_key(v) : ~0ull;
        }
        if (PF && item + nwarps < n) {
            const double *nrow = x + (item + nwarps) * pitch;
#pragma unroll
            for (int j = 0; j < FPL; ++j) { const int ff = lane + 32 * j; nv[j] = (ff < f) ? nrow[ff] : 0.0; }
        }
        const double med = warp_median_radix<FPL>(k, f, lane, hist);
        if (lane == 0) out_median[item] = med;
    }
}

// THR compute threads + AUXW auxiliary warps.  The left-to-right sums of phase A' are one dependent chain per item (the
// oracle's order), i.e. T busy threads for ~10 us per tile; they run on the auxiliary warps WHILE the compute warps walk
// the graph (phase B only needs the transposed tile), and meet them again at the reduction (phase C).
template <int R>
struct TmAux { static constexpr int T = 16 * R; static constexpr int WARPS = (T + 31) / 32; };

template <int FPL, int R, int CHN, int THR, bool PF>
__global__ void __launch_bounds__(THR + 32 * TmAux<R>::WARPS, 1)
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

    const bool is_aux = threadIdx.x >= THR;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = threadIdx.x & 15, p = threadIdx.x >> 4;
    const int64_t ntiles = (n + T - 1) / T;
    auto compute_sync = []() { asm volatile("bar.sync 1, %0;\n" ::"n"(THR) : "memory"); };   // the compute warps only

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t item0 = tile * T;
        __syncthreads();                                               // previous tile fully consumed

        // ---- A: load, transpose into shared memory (compute warps)
        if (!is_aux) {
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
        }
        __syncthreads();

        double en[R];
#pragma unroll
        for (int r = 0; r < R; ++r) en[r] = 0.0;
        if (is_aux) {
            // ---- A': left-to-right sums (norm^2; mean), one thread per item
            const int t = threadIdx.x - THR;
            if (t < T) {
                double n2 = 0.0, sm = 0.0;
#pragma unroll 8
                for (int ff = 0; ff < f; ++ff) {                       // loads and products run ahead; only the adds are a chain
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
        } else {
            // ---- B: x^T L x through the strictly-upper adjacency
            // PF (ASP_TM_PREFETCH=1, not yet the default: written after the GPU budget of round 1 was spent): the adjacency of
            // chunk c+1 is fetched into registers while chunk c is walked, so the L2 round trip of the staging -- 24 % of this
            // kernel's stall samples -- is no longer exposed between two barriers.
            constexpr int NPF = (CHN + THR - 1) / THR;
            int32_t pf_col[NPF], pf_rptr = 0;
            double pf_val[NPF], pf_deg = 0.0;
            int pf_ra0 = 0, pf_ra1 = 0, pf_e0 = 0, pf_e1 = 0;
            auto fetch = [&](int c) {
                pf_ra0 = chunks[c].row_begin; pf_ra1 = chunks[c].row_end;
                pf_e0 = uptr[pf_ra0]; pf_e1 = uptr[pf_ra1];
#pragma unroll
                for (int u = 0; u < NPF; ++u) {
                    const int i = (int)threadIdx.x + u * THR;
                    if (i < pf_e1 - pf_e0) { pf_col[u] = ucol[pf_e0 + i]; pf_val[u] = uval[pf_e0 + i]; }
                }
                if ((int)threadIdx.x <= pf_ra1 - pf_ra0) pf_rptr = uptr[pf_ra0 + threadIdx.x] - pf_e0;
                if ((int)threadIdx.x < pf_ra1 - pf_ra0) pf_deg = deg[pf_ra0 + threadIdx.x];
            };
            if (PF) fetch(0);
            for (int c = 0; c < nchunks; ++c) {
                int ra0, ra1;
                if (PF) {
                    ra0 = pf_ra0; ra1 = pf_ra1;
                    const int ne = pf_e1 - pf_e0;
                    compute_sync();                                    // the previous chunk is consumed
#pragma unroll
                    for (int u = 0; u < NPF; ++u) {
                        const int i = (int)threadIdx.x + u * THR;
                        if (i < ne) { s_col[i] = pf_col[u]; s_val[i] = pf_val[u]; }
                    }
                    if ((int)threadIdx.x <= ra1 - ra0) s_rptr[threadIdx.x] = pf_rptr;
                    if ((int)threadIdx.x < ra1 - ra0) s_deg[threadIdx.x] = pf_deg;
                    if (c + 1 < nchunks) fetch(c + 1);                 // in flight during the walk below
                } else {
                    ra0 = chunks[c].row_begin; ra1 = chunks[c].row_end;
                    const int e0 = uptr[ra0], e1 = uptr[ra1];
                    compute_sync();
                    for (int i = threadIdx.x; i < e1 - e0; i += THR) { s_col[i] = ucol[e0 + i]; s_val[i] = uval[e0 + i]; }
                    for (int i = threadIdx.x; i <= ra1 - ra0; i += THR) s_rptr[i] = uptr[ra0 + i] - e0;
                    for (int i = threadIdx.x; i < ra1 - ra0; i += THR) s_deg[i] = deg[ra0 + i];
                }
                compute_sync();
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
            compute_sync();                                            // the staged weights are dead: red[] aliases them
            // ---- C: the parts' partial energies
#pragma unroll
            for (int r = 0; r < R; ++r) red[p * T + g + 16 * r] = en[r];
        }
        __syncthreads();                                               // partial energies, norms and taus are in shared memory
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
        const char *mpf = getenv("ASP_MEDIAN_PREFETCH");
        auto mk = (mpf && mpf[0] == '1') ? median_kernel<FPL, true> : median_kernel<FPL, false>;
        mk<<<mgrid, 256, 0, ctx->stream>>>(x, n, f, pitch, sw->tau_mode == ASP_TAU_MEDIAN_ABS ? 1 : 0, medians);
        ASP_CUDA(cudaGetLastError());
        ASP_LAUNCHED(ctx);
    }
    static_assert((size_t)CHN * 12 >= (size_t)(THR / 16) * T * 8, "the part reduction aliases the staged weights and columns");
    const size_t smem = (size_t)f * (T + 1) * 8 + (size_t)CHN * 12 + (CH_ROWS + 2) * 4 + (size_t)CH_ROWS * 8 +
                        (size_t)T * 16 + 64;
    if (smem > 227 * 1024) ASP_FAIL(ASP_ERR_UNSUPPORTED, "taumode kernel: %d features do not fit in shared memory", f);
    // the register-prefetch variant needs one thread per staged row pointer (CH_ROWS + 1 <= THR)
    const char *pf_env = getenv("ASP_TM_PREFETCH");
    const bool pf = (THR > CH_ROWS) && pf_env && pf_env[0] == '1';
    auto kern = pf ? taumode_kernel<FPL, R, CHN, THR, (THR > CH_ROWS)> : taumode_kernel<FPL, R, CHN, THR, false>;
    ASP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;                                                       // resident CTAs per SM: their load / gather phases overlap
    constexpr int NTHREADS = THR + 32 * TmAux<R>::WARPS;
    ASP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NTHREADS, smem));
    if (occ < 1) occ = 1;
    const int64_t ntiles = (n + T - 1) / T;
    const int64_t slots = (int64_t)ctx->num_sms * occ;
    const int grid = (int)(ntiles < slots ? ntiles : slots);
    kern<<<grid, NTHREADS, smem, ctx->stream>>>(x, n, f, pitch, g->d_uptr, g->d_ucol, g->d_uval, g->d_deg, d_chunks,
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
