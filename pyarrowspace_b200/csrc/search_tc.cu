// search_tc.cu -- K4, throughput path: stage 1 of the batched lambda-aware search on the 5th-generation tensor
// cores (tcgen05.mma, accumulators in TMEM, operands by TMA), stage 2 exact.
// (replaces ArrowSpace::search_lambda_aware, /root/reference/src/lib.rs:173; score TAUMODE.md:33.)
//
// Why: the f64 score GEMM is bound by the FP64 tensor pipe (35 TFLOP/s measured); 2*Q*N*F = 1.26e13 FLOP per
// 16k-query step at C4 cannot go below ~360 ms there.  tcgen05 has no f64 kind, but the ANSWER only needs f64 on a
// few candidates per query: stage 1 computes every dot product from a two-term bf16 split
//     x = hi + lo + r,  |r| <= 2^-18 |x|        q.x ~ q_lo.x_hi + q_hi.x_lo + q_hi.x_hi      (3 MMAs, f32 accumulate)
// with a rigorous band  |cos~ - cos| <= DELTA_COS = 2^-13  (split truncation 3*2^-18 + f32 accumulation of 3F/16 MMA
// steps, 4x margin; checked against f64 in tests/test_gpu_parity.py::test_tc_dot_error_band), and EMITS every item
// whose approximate score could still be in the top-k:  s~ >= theta_k - 2 DELTA, theta_k = running k-th best.
// Stage 2 re-scores the emitted items in the reference order in f64 (exactly the oracle's expression) and selects
// top-k by (score desc, index asc).  Completeness: an item of the true top-k has s >= t_k >= theta* - DELTA, hence
// s~ >= theta* - 2 DELTA >= theta_run - 2 DELTA: it was emitted.  Only a full emission buffer sends a query to the
// exact scan.  The result is bit-identical to the FP64 path's (tests assert it).
//
// Kernel (one CTA per SM, 192 threads):
//   warp 0     TMA producer: 4-stage ring of {128 queries x 64 k, 256 items x 64 k} bf16 tiles, SWIZZLE_128B
//   warp 1     allocates TMEM (512 columns = 2 accumulators of 128 x 256 f32), issues tcgen05.mma (one lane),
//              tcgen05.commit -> frees the smem stage / publishes the accumulator
//   warps 2-5  epilogue: tcgen05.ld 32 lanes x 32 columns; thread <-> query row, so the running top-k
//              (16 registers) and the emission cursor are thread local: no atomics, no shuffles
// Bound: bf16 tensor pipe, 3 * 2*Q*N*Fp FLOP executed for 2*Q*N*F algorithmic.
#include "common.cuh"
#include "ptx.cuh"
#include "warp_sort.cuh"

#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

namespace {

using asp::Cand;

constexpr int TQ = 128;            // queries per CTA (UMMA M)
constexpr int TN = 256;            // items per tile (UMMA N)
constexpr int TKB = 64;            // bf16 per smem row = 128 B = one swizzle atom
constexpr int TC_STAGES = 4;
constexpr int A_BYTES = TQ * TKB * 2;          // 16 KB
constexpr int B_BYTES = TN * TKB * 2;          // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int TK_LIST = 16;        // running top list per (query row, column quarter) (topk <= 16)
constexpr int EPI_WARPS = 16;      // 4 per TMEM lane group: each thread owns one query row x 64 of the 256 tile columns
constexpr int EPI_SPLIT = EPI_WARPS / 4;
constexpr int EPI_COLS = TN / EPI_SPLIT;
constexpr int TC_THREADS = 64 + EPI_WARPS * 32;
constexpr float DELTA_COS = 1.220703125e-4f;   // 2^-13

// ------------------------------------------------------------------ f64 -> (hi, lo) bf16, f64 -> f32
// rows are scaled to unit length first (row_scale = 1/norm), so the tensor-core dot product IS the cosine and the
// epilogue compares raw accumulators against one per-row threshold
__global__ void split_bf16_kernel(const double *__restrict__ x, int64_t n, int f, int pitch, int kp,
                                  const double *__restrict__ row_scale, __nv_bfloat16 *__restrict__ hi,
                                  __nv_bfloat16 *__restrict__ lo)
{
    const int64_t total = n * (int64_t)kp;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / kp;
        const int c = (int)(i % kp);
        double v = (c < f) ? x[r * pitch + c] * row_scale[r] : 0.0;
        const __nv_bfloat16 h = __double2bfloat16(v);
        const double rem = v - (double)__bfloat162float(h);
        hi[i] = h;
        lo[i] = __double2bfloat16(rem);
    }
}

__global__ void to_f32_kernel(const double *__restrict__ a, int64_t n, float *__restrict__ fa)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        fa[i] = (float)a[i];
}

__global__ void minmax_kernel(const double *__restrict__ a, int64_t n, double *__restrict__ out /* [grid][2] */)
{
    double lo = INFINITY, hi = -INFINITY;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = a[i];
        lo = fmin(lo, v);
        hi = fmax(hi, v);
    }
    __shared__ double s_lo[256], s_hi[256];
    s_lo[threadIdx.x] = lo;
    s_hi[threadIdx.x] = hi;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) {
            s_lo[threadIdx.x] = fmin(s_lo[threadIdx.x], s_lo[threadIdx.x + off]);
            s_hi[threadIdx.x] = fmax(s_hi[threadIdx.x], s_hi[threadIdx.x + off]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[2 * blockIdx.x] = s_lo[0]; out[2 * blockIdx.x + 1] = s_hi[0]; }
}

// ------------------------------------------------------------------ descriptors
// K-major operand tile in shared memory, rows of 128 bytes, SWIZZLE_128B (what TMA wrote):
// 8-row groups are 1024 B apart (SBO), LBO unused (1), descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// kind::f16: D f32 (bit 4), A bf16 (bit 7), B bf16 (bit 10), both K-major, N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t IDESC_BF16_128x256 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TQ >> 4) << 24);

struct TcParams {
    int64_t nq, n_local;
    int kp;                       // padded feature count (multiple of 64)
    int nchunks;
    int capb;                     // emission capacity per (query, chunk)
    int topk;
    float tau, beta, delta;       // delta = band of one approximate score; tau > 0
    const float *lam_x, *lam_q;
    float lam_min, lam_max;       // range of the item lambdas of the shard (bounds the proximity term per query)
    float *emit_sc;
    int32_t *emit_ix;
    int32_t *emit_cnt;
    float *emit_theta;
    float *dump;                  // DUMP mode: raw dots [nq][n_local]
};

template <bool DUMP, int VARIANT>      // VARIANT (profiling only): 0 normal, 2 epilogue does no work, 3 no MMA issued
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_q_hi, const __grid_constant__ CUtensorMap map_q_lo,
               const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo, const TcParams p)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // SWIZZLE_128B tiles must start on 1024-byte boundaries
    unsigned char *stages = smem_raw + ((1024u - (asp::smem_u32(smem_raw) & 1023u)) & 1023u);   // TC_STAGES * STAGE_BYTES
    float *s_const = reinterpret_cast<float *>(stages + TC_STAGES * STAGE_BYTES);   // [2][TN]: item lambdas per accumulator
    __shared__ __align__(8) uint64_t full_bar[TC_STAGES], empty_bar[TC_STAGES], tmem_full[2], tmem_empty[2];
    __shared__ uint32_t s_tmem_base;
    __shared__ uint32_t s_theta[TQ];          // per query row: best k-th score any of its EPI_SPLIT threads has seen (ordered bits)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qb = blockIdx.x, chunk = blockIdx.y;
    const int64_t tiles_total = (p.n_local + TN - 1) / TN;
    const int64_t tile0 = (tiles_total * chunk) / p.nchunks;
    const int64_t ntiles = (tiles_total * (chunk + 1)) / p.nchunks - tile0;
    const int ksteps = p.kp / TKB;
    const int kiters = 3 * ksteps;                                               // 3 split terms

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { asp::mbar_init(&full_bar[s], 1); asp::mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { asp::mbar_init(&tmem_full[a], 1); asp::mbar_init(&tmem_empty[a], EPI_WARPS); }
        asp::fence_barrier_init();
    }
    if (threadIdx.x < TQ) s_theta[threadIdx.x] = 0u;                             // ordered bits of -inf... (0 = below every float)
    if (warp == 1) asp::tmem_alloc(&s_tmem_base, 512);
    asp::tc_fence_before();
    __syncthreads();
    asp::tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            asp::tma_prefetch_desc(&map_q_hi); asp::tma_prefetch_desc(&map_q_lo);
            asp::tma_prefetch_desc(&map_x_hi); asp::tma_prefetch_desc(&map_x_lo);
            int64_t it = 0;
            for (int64_t t = 0; t < ntiles; ++t) {
                const int item0 = (int)((tile0 + t) * TN);
                for (int ki = 0; ki < kiters; ++ki, ++it) {
                    const int s = (int)(it % TC_STAGES);
                    asp::mbar_wait(&empty_bar[s], (uint32_t)(((it / TC_STAGES) & 1) ^ 1));
                    const int seg = ki / ksteps, kc = (ki % ksteps) * TKB;
                    // small terms first: (q_lo, x_hi), (q_hi, x_lo), then (q_hi, x_hi)
                    const CUtensorMap *ma = (seg == 0) ? &map_q_lo : &map_q_hi;
                    const CUtensorMap *mb = (seg == 1) ? &map_x_lo : &map_x_hi;
                    unsigned char *dst = stages + (size_t)s * STAGE_BYTES;
                    asp::mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
                    asp::tma_load_2d(dst, ma, &full_bar[s], kc, qb * TQ);
                    asp::tma_load_2d(dst + A_BYTES, mb, &full_bar[s], kc, item0);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int64_t it = 0;
            for (int64_t t = 0; t < ntiles; ++t) {
                const int acc = (int)(t & 1);
                asp::mbar_wait(&tmem_empty[acc], (uint32_t)((((t >> 1) & 1)) ^ 1));
                asp::tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * TN);
                for (int ki = 0; ki < kiters; ++ki, ++it) {
                    const int s = (int)(it % TC_STAGES);
                    asp::mbar_wait(&full_bar[s], (uint32_t)((it / TC_STAGES) & 1));
                    asp::tc_fence_after();
                    const uint32_t a_addr = asp::smem_u32(stages + (size_t)s * STAGE_BYTES);
                    const uint64_t da = make_kmajor_sw128_desc(a_addr);
                    const uint64_t db = make_kmajor_sw128_desc(a_addr + A_BYTES);
                    if (VARIANT != 3) {
#pragma unroll
                        for (int k = 0; k < TKB / 16; ++k)                        // UMMA K = 16 bf16 = 32 B = +2 in the address field
                            asp::umma_f16(tmem_d, da + 2 * k, db + 2 * k, IDESC_BF16_128x256, (ki > 0 || k > 0) ? 1u : 0u);
                    }
                    asp::umma_commit(&empty_bar[s]);                             // smem stage reusable when these MMAs retire
                }
                asp::umma_commit(&tmem_full[acc]);                               // accumulator complete
            }
        }
    } else {
        // ===================== epilogue: warps 2..5, thread <-> query row =====================
        const int lg = warp & 3;                                                 // TMEM lane group this warp may touch
        const int part = (warp - 2) >> 2;                                        // which EPI_COLS columns of the tile
        const int row = lg * 32 + lane;
        const int64_t gq = (int64_t)qb * TQ + row;
        const bool qvalid = gq < p.nq;
        const float lq = qvalid ? p.lam_q[gq] : 0.f;
        const int et = threadIdx.x - 64;                                         // 0 .. EPI_WARPS*32-1
        // score <= tau*cos + beta*prox_ub, prox_ub = 1/(1 + distance of lambda_q to the shard's lambda range)
        // (beta >= 0 on this path; the 0.999 keeps the bound safe against the f32 roundings of the range)
        const float lam_gap = fmaxf(0.f, fmaxf(lq - p.lam_max, p.lam_min - lq)) * 0.999f;
        const float beta_ub = p.beta * __fdividef(1.0f, 1.0f + lam_gap) * 1.000001f;
        const float inv_tau = 1.0f / p.tau;
        float lst[TK_LIST];
#pragma unroll
        for (int i = 0; i < TK_LIST; ++i) lst[i] = -INFINITY;
        float theta_k = -INFINITY, theta_emit = -INFINITY;
        float theta_dot = qvalid ? -INFINITY : INFINITY;                         // raw-accumulator (cosine) filter
        int cnt = 0;
        const size_t ebase = (((size_t)gq * p.nchunks + chunk) * EPI_SPLIT + part) * (size_t)p.capb;
        // order-preserving float <-> uint32 (so the row threshold can be shared with atomicMax)
        auto f2o = [](float f) { const uint32_t b = __float_as_uint(f); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); };
        auto o2f = [](uint32_t o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o); };

        for (int64_t t = 0; t < ntiles; ++t) {
            const int acc = (int)(t & 1);
            const int64_t item0 = (tile0 + t) * TN;
            float *c_lam = s_const + acc * TN;
            for (int j = et; j < TN; j += EPI_WARPS * 32) {
                const int64_t n = item0 + j;
                c_lam[j] = (n < p.n_local) ? p.lam_x[n] : 0.f;
            }
            asm volatile("bar.sync 1, %0;\n" ::"n"(EPI_WARPS * 32) : "memory");    // epilogue warps only
            if (!DUMP && qvalid) {                                               // adopt the row's shared threshold
                const uint32_t so = s_theta[row];
                if (so != 0u) {
                    const float sh = o2f(so);
                    if (sh > theta_k) {
                        theta_k = sh;
                        theta_emit = theta_k - 2.0f * p.delta;
                        theta_dot = (theta_emit - beta_ub) * inv_tau - 1e-6f;
                    }
                }
            }
            asp::mbar_wait(&tmem_full[acc], (uint32_t)((t >> 1) & 1));
            asp::tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < (VARIANT >= 2 ? 0 : EPI_COLS / 32); ++c) {
                const int col0 = part * EPI_COLS + c * 32;
                uint32_t r[32];
                asp::tmem_ld_32x32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * TN + col0), r);
                asp::tmem_ld_wait();
                if (DUMP) {
                    if (qvalid) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int64_t n = item0 + col0 + j;
                            if (n < p.n_local) p.dump[gq * p.n_local + n] = __uint_as_float(r[j]);
                        }
                    }
                } else {
                    // common path: the accumulator IS the cosine (unit operands); a max tree and one compare
                    float m[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) m[j] = fmaxf(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
#pragma unroll
                    for (int w = 8; w > 0; w >>= 1)
#pragma unroll
                        for (int j = 0; j < w; ++j) m[j] = fmaxf(m[j], m[j + w]);
                    if (m[0] >= theta_dot) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float d = __uint_as_float(r[j]);
                            if (d >= theta_dot) {
                                const int col = col0 + j;
                                const int64_t n = item0 + col;
                                const float sc = fmaf(p.beta, __fdividef(1.0f, 1.0f + fabsf(lq - c_lam[col])), p.tau * d);
                                if (sc >= theta_emit && n < p.n_local) {
                                    if (cnt < p.capb) { p.emit_sc[ebase + cnt] = sc; p.emit_ix[ebase + cnt] = (int32_t)n; }
                                    ++cnt;
                                    if (sc > lst[TK_LIST - 1]) {
                                        float v = sc;
#pragma unroll
                                        for (int i = 0; i < TK_LIST; ++i) {
                                            const float o = lst[i];
                                            const bool sw = v > o;
                                            lst[i] = sw ? v : o;
                                            v = sw ? o : v;
                                        }
                                        float th = lst[0];
#pragma unroll
                                        for (int i = 1; i < TK_LIST; ++i) th = (i < p.topk) ? lst[i] : th;
                                        if (th > theta_k) {
                                            theta_k = th;
                                            atomicMax(&s_theta[row], f2o(th));
                                            theta_emit = theta_k - 2.0f * p.delta;
                                            // s <= tau*cos + beta*prox_ub: below this cosine nothing can reach theta_emit
                                            theta_dot = (theta_emit - beta_ub) * inv_tau - 1e-6f;
                                        }
                                    }
                                }
                            }
                        }
                    }
                }
            }
            asp::tc_fence_before();
            __syncwarp();
            if (lane == 0) asp::mbar_arrive(&tmem_empty[acc]);
        }
        if (!DUMP && qvalid) {
            p.emit_cnt[(gq * p.nchunks + chunk) * EPI_SPLIT + part] = cnt;
            p.emit_theta[(gq * p.nchunks + chunk) * EPI_SPLIT + part] = theta_k;
        }
    }
    asp::tc_fence_before();
    __syncthreads();
    if (warp == 1) asp::tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------ stage 2
__device__ __forceinline__ double exact_score_tc(double dot, double nq, double nx, double tau, double lq, double lx)
{
    const double den = __dmul_rn(nq, nx);
    const double c = (den == 0.0) ? 0.0 : __ddiv_rn(dot, den);
    const double prox = __ddiv_rn(1.0, __dadd_rn(1.0, fabs(__dsub_rn(lq, lx))));
    return __dadd_rn(__dmul_rn(tau, c), __dmul_rn(__dsub_rn(1.0, tau), prox));
}

__device__ __forceinline__ double seq_dot_row_tc(const double *__restrict__ qv, const double *__restrict__ row, int f)
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

constexpr int TR_WARPS = 4;

// One warp per query: filter the emitted candidates with the final threshold, re-score the survivors in f64 in the
// reference order (one lane per candidate), keep the best 32 by (score desc, index asc), emit top-k.
__global__ void __launch_bounds__(TR_WARPS * 32)
tc_rescore_kernel(const double *__restrict__ q, int qpitch, int64_t nq, const double *__restrict__ items, int64_t n_local,
                  int f, int pitch, int64_t row0, const double *__restrict__ norm_x, const double *__restrict__ lam_x,
                  const double *__restrict__ norm_q, const double *__restrict__ lam_q, double tau, int topk, int nchunks,
                  int capb, float delta, const float *__restrict__ emit_sc, const int32_t *__restrict__ emit_ix,
                  const int32_t *__restrict__ emit_cnt, const float *__restrict__ emit_theta,
                  int64_t *__restrict__ out_idx, double *__restrict__ out_score, int32_t *slow_list, int32_t *slow_count,
                  unsigned long long *survivor_total)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *qs = reinterpret_cast<double *>(smem_raw) + (size_t)warp * f;
    int32_t *queue = reinterpret_cast<int32_t *>(reinterpret_cast<double *>(smem_raw) + (size_t)TR_WARPS * f) + warp * 64;
    const int64_t qi = (int64_t)blockIdx.x * TR_WARPS + warp;
    if (qi >= nq) return;
    for (int j = lane; j < f; j += 32) qs[j] = q[qi * qpitch + j];

    float theta = -INFINITY;
    bool overflow = false;
    for (int c = lane; c < nchunks; c += 32) {
        theta = fmaxf(theta, emit_theta[qi * nchunks + c]);
        overflow |= emit_cnt[qi * nchunks + c] > capb;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) theta = fmaxf(theta, __shfl_xor_sync(0xffffffffu, theta, off));
    if (__any_sync(0xffffffffu, overflow)) {
        if (lane == 0) slow_list[atomicAdd(slow_count, 1)] = (int32_t)qi;
        return;
    }
    const float cutoff = theta - 2.0f * delta;
    const double nqv = norm_q[qi], lqv = lam_q[qi];
    __syncwarp();

    Cand best[2];
    best[0] = asp::cand_empty();
    best[1] = asp::cand_empty();
    int qn = 0;                                                                  // queued survivors (warp uniform)
    unsigned long long nsurv = 0;
    auto flush = [&](int count) {
        // lanes < count re-score one survivor each; merged into the running best 32
        Cand c = asp::cand_empty();
        if (lane < count) {
            const int64_t it = queue[lane];
            const double d = seq_dot_row_tc(qs, items + it * pitch, f);
            c.s = exact_score_tc(d, nqv, norm_x[it], tau, lqv, lam_x[it]);
            c.i = (int32_t)it;
        }
        best[1] = c;
        asp::warp_sort_best_first<2>(best, lane);
    };
    for (int c = 0; c < nchunks; ++c) {
        const int cnt = emit_cnt[qi * nchunks + c];
        const size_t base = ((size_t)qi * nchunks + c) * (size_t)capb;
        for (int e0 = 0; e0 < cnt; e0 += 32) {
            const int e = e0 + lane;
            const bool keep = (e < cnt) && (emit_sc[base + e] >= cutoff);
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) queue[qn + __popc(m & ((1u << lane) - 1))] = emit_ix[base + e];
            qn += __popc(m);
            __syncwarp();
            if (qn >= 32) {
                flush(32);
                nsurv += 32;
                __syncwarp();
                if (lane < qn - 32) queue[lane] = queue[32 + lane];
                qn -= 32;
                __syncwarp();
            }
        }
    }
    if (qn > 0) { flush(qn); nsurv += qn; }
    if (lane == 0) atomicAdd(survivor_total, nsurv);
    const int kk = topk < n_local ? topk : (int)n_local;
    if (lane < topk) {
        const bool ok = (lane < kk) && best[0].i != 0x7fffffff;
        out_idx[qi * topk + lane] = ok ? row0 + best[0].i : -1;
        out_score[qi * topk + lane] = ok ? best[0].s : NAN;
    }
}

}  // namespace

// ------------------------------------------------------------------ host side
struct asp_tc_cache {               // per-space bf16 copies, built on the first tensor-core search
    __nv_bfloat16 *hi = nullptr, *lo = nullptr;
    float *lam32 = nullptr;
    float lam_min = 0.f, lam_max = 0.f;
    int kp = 0;
    CUtensorMap map_hi, map_lo;
};

static int ensure_tc_cache(const asp_space *s, asp_tc_cache **out)
{
    asp_space *ms = const_cast<asp_space *>(s);
    if (ms->tc_cache) { *out = static_cast<asp_tc_cache *>(ms->tc_cache); return ASP_OK; }
    asp_ctx *ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    asp_tc_cache *c = new asp_tc_cache();
    c->kp = (s->f + TKB - 1) / TKB * TKB;
    const size_t ne = (size_t)s->n_local * c->kp;
    ASP_CUDA(cudaMallocAsync(&c->hi, ne * 2, st));
    ASP_CUDA(cudaMallocAsync(&c->lo, ne * 2, st));
    ASP_CUDA(cudaMallocAsync(&c->lam32, sizeof(float) * s->n_local, st));
    split_bf16_kernel<<<ctx->num_sms * 8, 256, 0, st>>>(s->items, s->n_local, s->f, s->fp, c->kp, s->inv_norms, c->hi, c->lo);
    ASP_CUDA(cudaGetLastError()); ASP_LAUNCHED(ctx);
    to_f32_kernel<<<ctx->num_sms * 2, 256, 0, st>>>(s->lambdas, s->n_local, c->lam32);
    ASP_CUDA(cudaGetLastError()); ASP_LAUNCHED(ctx);
    {
        double *d_mm = nullptr;
        std::vector<double> h_mm(2 * 256);
        ASP_CUDA(cudaMallocAsync(&d_mm, sizeof(double) * 512, st));
        minmax_kernel<<<256, 256, 0, st>>>(s->lambdas, s->n_local, d_mm);
        ASP_CUDA(cudaGetLastError()); ASP_LAUNCHED(ctx);
        ASP_CUDA(cudaMemcpyAsync(h_mm.data(), d_mm, sizeof(double) * 512, cudaMemcpyDeviceToHost, st));
        ASP_CUDA(cudaStreamSynchronize(st));
        ASP_CUDA(cudaFreeAsync(d_mm, st));
        double lo = INFINITY, hi = -INFINITY;
        for (int i = 0; i < 256; ++i) { lo = fmin(lo, h_mm[2 * i]); hi = fmax(hi, h_mm[2 * i + 1]); }
        c->lam_min = (float)lo; c->lam_max = (float)hi;
        if ((double)c->lam_min > lo) c->lam_min = nextafterf(c->lam_min, -INFINITY);   // round outwards
        if ((double)c->lam_max < hi) c->lam_max = nextafterf(c->lam_max, INFINITY);
    }
    ASP_CHECK(asp_make_bf16_tmap(&c->map_hi, c->hi, s->n_local, c->kp, TN));
    ASP_CHECK(asp_make_bf16_tmap(&c->map_lo, c->lo, s->n_local, c->kp, TN));
    ms->tc_cache = c;
    *out = c;
    return ASP_OK;
}

void asp_free_tc_cache(asp_space *s)
{
    if (!s->tc_cache) return;
    asp_tc_cache *c = static_cast<asp_tc_cache *>(s->tc_cache);
    cudaStream_t st = s->ctx->stream;
    cudaFreeAsync(c->hi, st); cudaFreeAsync(c->lo, st); cudaFreeAsync(c->lam32, st);
    delete c;
    s->tc_cache = nullptr;
}

bool asp_search_tc_supported(const asp_space *s, int64_t nq, int64_t topk, double tau)
{
    return topk >= 1 && topk <= TK_LIST && nq >= 1 && s->n_local >= 1 && s->n_local < 2147483647LL && tau > 1e-3 &&
           tau <= 1.0;            // beta = 1 - tau >= 0: the proximity term is bounded from above by prox_ub
}

// dump == nullptr: full search.  dump != nullptr: raw approximate dots [nq][n_local] f32 (tests).
int asp_search_tc_impl(const asp_space *s, const double *q_dev, int64_t nq, int32_t qpitch, const double *lambda_q_dev,
                       const double *qnorm_dev, double tau, int64_t topk, int64_t *out_idx_dev, double *out_score_dev,
                       float *dump_dev)
{
    asp_ctx *ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    asp_tc_cache *c = nullptr;
    ASP_CHECK(ensure_tc_cache(s, &c));
    const int kp = c->kp;

    // queries: bf16 split + f32 scalars
    __nv_bfloat16 *q_hi = nullptr, *q_lo = nullptr;
    float *lam_q32 = nullptr;
    double *inv_nq = nullptr;
    ASP_CUDA(cudaMallocAsync(&q_hi, (size_t)nq * kp * 2, st));
    ASP_CUDA(cudaMallocAsync(&q_lo, (size_t)nq * kp * 2, st));
    ASP_CUDA(cudaMallocAsync(&lam_q32, sizeof(float) * nq, st));
    ASP_CUDA(cudaMallocAsync(&inv_nq, sizeof(double) * nq, st));
    ASP_CHECK(asp_launch_reciprocal(ctx, qnorm_dev, nq, inv_nq));
    split_bf16_kernel<<<ctx->num_sms * 4, 256, 0, st>>>(q_dev, nq, s->f, qpitch, kp, inv_nq, q_hi, q_lo);
    ASP_CUDA(cudaGetLastError()); ASP_LAUNCHED(ctx);
    to_f32_kernel<<<64, 256, 0, st>>>(lambda_q_dev, nq, lam_q32);
    ASP_CUDA(cudaGetLastError()); ASP_LAUNCHED(ctx);
    CUtensorMap map_q_hi, map_q_lo;
    ASP_CHECK(asp_make_bf16_tmap(&map_q_hi, q_hi, nq, kp, TQ));
    ASP_CHECK(asp_make_bf16_tmap(&map_q_lo, q_lo, nq, kp, TQ));

    // grid: query blocks x item chunks, whole waves of SMs
    const int64_t tiles_total = asp_ceil_div(s->n_local, TN);
    const int64_t qblocks = asp_ceil_div(nq, TQ);
    int64_t best_chunks = 1;
    double best_eff = 0.0;
    for (int w = 1; w <= 16; ++w) {
        int64_t cc = ((int64_t)w * ctx->num_sms) / qblocks;
        if (cc < 1) continue;
        if (cc > tiles_total) cc = tiles_total;
        const int64_t ctas = cc * qblocks;
        const double eff = (double)ctas / (double)(asp_ceil_div(ctas, ctx->num_sms) * ctx->num_sms);
        if (eff > best_eff + 1e-9) { best_eff = eff; best_chunks = cc; }
        if (eff >= 0.97 || cc == tiles_total) break;
    }
    const int nchunks = (int)best_chunks;
    int capb = dump_dev ? 1 : 1024 / EPI_SPLIT;                      // per (query, chunk, column part)
    if (const char *e = getenv("ASP_TC_CAPB")) {                     // test knob: shrink the emission buffers
        const int v = atoi(e);
        if (!dump_dev && v >= 8 && v <= 65536) capb = v;
    }
    const int nsub = nchunks * EPI_SPLIT;

    TcParams p;
    p.nq = nq; p.n_local = s->n_local; p.kp = kp; p.nchunks = nchunks; p.capb = capb; p.topk = (int)topk;
    p.tau = (float)tau; p.beta = (float)(1.0 - tau);
    p.delta = (float)(fabs(tau) * DELTA_COS + (fabs(tau) + fabs(1.0 - tau)) * 2e-6);
    p.lam_x = c->lam32; p.lam_q = lam_q32; p.lam_min = c->lam_min; p.lam_max = c->lam_max;
    p.emit_sc = nullptr; p.emit_ix = nullptr; p.emit_cnt = nullptr; p.emit_theta = nullptr; p.dump = dump_dev;
    int32_t *slow_list = nullptr, *slow_count = nullptr;
    unsigned long long *survivors = nullptr;
    if (!dump_dev) {
        ASP_CUDA(cudaMallocAsync(&p.emit_sc, sizeof(float) * (size_t)nq * nsub * capb, st));
        ASP_CUDA(cudaMallocAsync(&p.emit_ix, sizeof(int32_t) * (size_t)nq * nsub * capb, st));
        ASP_CUDA(cudaMallocAsync(&p.emit_cnt, sizeof(int32_t) * (size_t)nq * nsub, st));
        ASP_CUDA(cudaMallocAsync(&p.emit_theta, sizeof(float) * (size_t)nq * nsub, st));
        ASP_CUDA(cudaMallocAsync(&slow_list, sizeof(int32_t) * (nq + 1), st));
        ASP_CUDA(cudaMallocAsync(&slow_count, sizeof(int32_t), st));
        ASP_CUDA(cudaMallocAsync(&survivors, sizeof(unsigned long long), st));
        ASP_CUDA(cudaMemsetAsync(slow_count, 0, sizeof(int32_t), st));
        ASP_CUDA(cudaMemsetAsync(survivors, 0, sizeof(unsigned long long), st));
    }

    const size_t smem = (size_t)TC_STAGES * STAGE_BYTES + 2 * TN * sizeof(float) + 1024;
    dim3 grid((unsigned)qblocks, nchunks);
    ASP_CUDA(cudaEventRecord(ctx->ev0, st));
    if (dump_dev) {
        ASP_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_gemm_kernel<true, 0><<<grid, TC_THREADS, smem, st>>>(map_q_hi, map_q_lo, c->map_hi, c->map_lo, p);
    } else {
        const char *var = getenv("ASP_TC_VARIANT");               // profiling only: results are wrong for 2 / 3
        const int v = var ? atoi(var) : 0;
        auto k = (v == 2) ? tc_gemm_kernel<false, 2> : (v == 3) ? tc_gemm_kernel<false, 3> : tc_gemm_kernel<false, 0>;
        ASP_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, TC_THREADS, smem, st>>>(map_q_hi, map_q_lo, c->map_hi, c->map_lo, p);
    }
    ASP_CUDA(cudaGetLastError()); ASP_LAUNCHED(ctx);
    ASP_CUDA(cudaEventRecord(ctx->ev1, st));

    int rc = ASP_OK;
    if (!dump_dev) {
        const size_t rsmem = (size_t)TR_WARPS * s->f * 8 + TR_WARPS * 64 * 4;
        ASP_CUDA(cudaFuncSetAttribute(tc_rescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem));
        tc_rescore_kernel<<<(unsigned)asp_ceil_div(nq, TR_WARPS), TR_WARPS * 32, rsmem, st>>>(
            q_dev, qpitch, nq, s->items, s->n_local, s->f, s->fp, s->row0, s->norms, s->lambdas, qnorm_dev, lambda_q_dev, tau,
            (int)topk, nsub, capb, p.delta, p.emit_sc, p.emit_ix, p.emit_cnt, p.emit_theta, out_idx_dev, out_score_dev,
            slow_list, slow_count, survivors);
        ASP_CUDA(cudaGetLastError()); ASP_LAUNCHED(ctx);
        int32_t nslow = 0;
        unsigned long long nsurv = 0;
        ASP_CUDA(cudaMemcpyAsync(&nslow, slow_count, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        ASP_CUDA(cudaMemcpyAsync(&nsurv, survivors, sizeof(nsurv), cudaMemcpyDeviceToHost, st));
        ASP_CUDA(cudaStreamSynchronize(st));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        ctx->stats["search_stage1_ms"] = ms;
        ctx->stats["search_slow_queries"] = nslow;
        ctx->stats["search_rescored_per_query"] = (double)nsurv / (double)nq;
        ctx->stats["search_stage1_is_tc"] = 1.0;
        if (nslow > 0)
            rc = asp_search_slow_path(s, q_dev, qpitch, lambda_q_dev, qnorm_dev, tau, topk, slow_list, nslow, out_idx_dev,
                                      out_score_dev);
        cudaFreeAsync(p.emit_sc, st); cudaFreeAsync(p.emit_ix, st); cudaFreeAsync(p.emit_cnt, st);
        cudaFreeAsync(p.emit_theta, st); cudaFreeAsync(slow_list, st); cudaFreeAsync(slow_count, st);
        cudaFreeAsync(survivors, st);
    }
    cudaFreeAsync(q_hi, st); cudaFreeAsync(q_lo, st); cudaFreeAsync(lam_q32, st);
    cudaFreeAsync(inv_nq, st);
    return rc;
}
