// common.cuh -- shared declarations of the sm_100a implementation of the arrowspace hot path.
#pragma once

#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <functional>
#include <map>
#include <vector>

#include "../../include/arrowspace_b200.h"

// ------------------------------------------------------------------ errors
void asp_set_error(const char *fmt, ...);

#define ASP_CUDA(call)                                                                       \
    do {                                                                                     \
        cudaError_t _e = (call);                                                             \
        if (_e != cudaSuccess) {                                                             \
            asp_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__,      \
                          __LINE__, cudaGetErrorString(_e));                                 \
            return ASP_ERR_CUDA;                                                             \
        }                                                                                    \
    } while (0)

#define ASP_CHECK(expr)                                                                      \
    do {                                                                                     \
        int _rc = (expr);                                                                    \
        if (_rc != ASP_OK) return _rc;                                                       \
    } while (0)

#define ASP_FAIL(code, ...)                                                                  \
    do {                                                                                     \
        asp_set_error(__VA_ARGS__);                                                          \
        return (code);                                                                       \
    } while (0)

// ------------------------------------------------------------------ handles
struct asp_ctx {
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    int64_t launches = 0;
    bool use_tma = true;                 // ASP_NO_TMA=1 switches the operand loaders to cp.async
    std::map<std::string, double> stats;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    cudaStream_t up_stream = nullptr, down_stream = nullptr;   // copy streams of the pipelined host search (lazily created)
    cudaEvent_t pipe_ev[7] = {};         // its events (uploads, piece done, piece start, allocation ready), created with the streams
    std::function<void()> on_wait;       // host work the search runs once, right before its long wait for the kernels
};

struct asp_space {
    asp_ctx *ctx = nullptr;
    int64_t n_local = 0, row0 = 0, n_total = 0;
    int32_t f = 0;
    int32_t fp = 0;                      // row pitch in doubles: f rounded up to a multiple of 4, zero padded
    double *items = nullptr;             // device, n_local x fp
    bool owns_items = true;              // false: adopted from the caller (asp_space_adopt)
    double *norms = nullptr;             // device, n_local: sqrt(sum x^2), left-to-right
    double *inv_norms = nullptr;         // device, n_local: 1/norm (0 for zero rows)
    double *lambdas = nullptr;           // device, n_local
    bool have_lambdas = false;
    CUtensorMap tmap_gram;               // (4, n_local, fp/4), box (4, 32 rows, 32 quads): Gram kernel
    CUtensorMap tmap_rows;               // same tensor, box (4, 128 rows, 4 quads): search / kNN GEMM
    int world = 1, rank = 0;             // shard = Gram segments [rank*8/world, (rank+1)*8/world)
    void *tc_cache = nullptr;            // fp16 split copies (lambda order) for the tcgen05 search path (search_tc.cu), lazily built
};

struct asp_graph {
    asp_ctx *ctx = nullptr;
    int64_t nnodes = 0, nnz = 0;
    asp_graph_params gp;
    asp_switches sw;
    // Laplacian CSR, device + host mirror
    int64_t *d_indptr = nullptr;
    int32_t *d_indices = nullptr;
    double  *d_data = nullptr;
    std::vector<int64_t> h_indptr;
    std::vector<int32_t> h_indices;
    std::vector<double>  h_data;
    // graph-only inputs of the lambda pass (taumode.cu: the symmetrised quadratic form, pre-cut into chunks); feature graphs only
    void *tm_blob = nullptr;
};

// host wall clock in microseconds (diagnostic stats "search_host_*_us": where a call's time goes between the kernels)
#include <chrono>
static inline double asp_now_us()
{
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ------------------------------------------------------------------ launch bookkeeping
#define ASP_LAUNCHED(ctx) ((ctx)->launches++)

static inline int64_t asp_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// neighbour cap of one node (SURVEY.md Appendix A4 + the unpinned conventions): min(k, topk) when topk prunes, minus the
// node itself when k counts it; never more than m - 1 others
static inline int64_t asp_neighbour_cap(const asp_graph_params *gp, const asp_switches *sw, int64_t m)
{
    int64_t kk = gp->k;
    if (sw && sw->topk_prunes && gp->topk < kk) kk = gp->topk;
    if (sw && sw->k_counts_self) kk -= 1;
    if (kk > m - 1) kk = m - 1;
    return kk < 0 ? 0 : kk;
}

// host helpers implemented in api.cu
int  asp_copy_in(asp_ctx *ctx, void *dst_dev, const void *src, size_t bytes);
int  asp_copy_out(asp_ctx *ctx, void *dst, const void *src_dev, size_t bytes);
bool asp_is_device_ptr(const void *p);

// tensormap.cu
int asp_make_items_tmap(CUtensorMap *out, const double *base, int64_t rows, int32_t fp,
                        int box_rows, int box_outer);

// gram.cu
int asp_launch_gram_partials(asp_space *s, double *out_dev);
int asp_launch_gram_reduce(asp_ctx *ctx, const double *segments_dev, int32_t f, double *gram_dev);

// graph_select.cu
struct asp_knn_lists {                    // device buffers, m rows x kk slots
    int64_t m = 0; int32_t kk = 0;
    int32_t *idx = nullptr; double *dist = nullptr; int32_t *cnt = nullptr;
};
int asp_feature_select(asp_ctx *ctx, const double *gram_dev, int32_t f, int64_t n_total,
                       const asp_graph_params *gp, const asp_switches *sw, const int32_t *exact_pairs_dev,
                       const double *exact_sums_dev, int64_t n_exact, asp_knn_lists *lists,
                       int32_t *need_pairs_dev, int64_t need_cap, int32_t *need_count_dev);
int asp_launch_exact_pairs(asp_space *s, const int32_t *pairs_dev, int64_t n_pairs, double *sums_dev);

// csr.cu
int asp_graph_host_mirror(asp_graph *g);
int asp_assemble_laplacian(asp_ctx *ctx, const asp_knn_lists *lists, const asp_graph_params *gp,
                           const asp_switches *sw, asp_graph *g);

// taumode.cu
int asp_graph_upload_upper(asp_graph *g);
void asp_graph_free_upper(asp_graph *g);
int asp_launch_taumode(asp_ctx *ctx, const asp_graph *g, const asp_switches *sw, const double *x_dev,
                       int64_t n, int32_t f, int32_t pitch, double *out_energy, double *out_tau,
                       double *out_lambda, double *out_norm, double *out_inv_norm, int *zero_flag_dev);

// search.cu
int asp_search_impl(const asp_space *s, const asp_graph *g, const double *q_dev, int64_t nq, int32_t qpitch,
                    const double *lambda_q_dev, const double *qnorm_dev, double tau, int64_t topk,
                    int64_t *out_idx_dev, double *out_score_dev);
int asp_topk_merge_impl(asp_ctx *ctx, const int64_t *idx_dev, const double *score_dev, int parts,
                        int64_t nq, int64_t topk, int64_t *out_idx_dev, double *out_score_dev);

// search_tc.cu (tcgen05 stage 1)
struct asp_tc_batch {                     // device buffers of one batch of query vectors between stage 1 and stage 2
    int64_t nq = 0;
    int nsub = 0, capb = 0, nterms = 0, variant = 0;
    double delta_cos_max = 0.0;
    void *q_hi = nullptr, *q_lo = nullptr;
    float *lam_q32 = nullptr, *delta_q = nullptr;    // visiting order
    double *inv_nq = nullptr, *rho_q = nullptr;
    int32_t *qperm = nullptr, *center = nullptr;     // visiting position -> caller's query index (nullptr: identity)
    float *emit_sc = nullptr;                        // [nq][nsub][capb] approximate scores
    int32_t *emit_ix = nullptr, *emit_cnt = nullptr; // local item indices; [nq][nsub] counts (> capb: overflow)
    uint32_t *theta_glob = nullptr;
};
int asp_tc_stage1(const asp_space *s, const double *q_dev, int64_t nq, int32_t qpitch, const double *lambda_q_dev,
                  const double *qnorm_dev, double tau, int64_t topk, double score_floor, int force_terms, int capb_override,
                  float *dump_dev, asp_tc_batch *b);
void asp_tc_batch_free(asp_ctx *ctx, asp_tc_batch *b);
bool asp_search_tc_supported(const asp_space *s, int64_t nq, int64_t topk, double tau);
int asp_search_tc_impl(const asp_space *s, const double *q_dev, int64_t nq, int32_t qpitch, const double *lambda_q_dev,
                       const double *qnorm_dev, double tau, int64_t topk, int64_t *out_idx_dev, double *out_score_dev,
                       float *dump_dev);
void asp_free_tc_cache(asp_space *s);
int asp_launch_reciprocal(asp_ctx *ctx, const double *x, int64_t n, double *out);
int asp_search_slow_path(const asp_space *s, const double *q_dev, int32_t qpitch, const double *lambda_q_dev,
                         const double *qnorm_dev, double tau, int64_t topk, const int32_t *slow_list_dev, int nslow,
                         int64_t *out_idx_dev, double *out_score_dev);
int asp_make_f16_tmap(CUtensorMap *out, const void *base, int64_t rows, int32_t kp, int box_rows);
int asp_make_f16_tmap_narrow(CUtensorMap *out, const void *base, int64_t rows, int32_t kp, int box_rows, int nw);

// knn.cu
int asp_item_knn(asp_space *s, const asp_graph_params *gp, asp_knn_lists *lists);
int asp_item_knn_rows_impl(asp_space *s, const asp_graph_params *gp, int64_t row_begin, int64_t row_end, asp_knn_lists *lists);
