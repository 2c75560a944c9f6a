/*
 * oracle.c -- CPU restatement of the pyarrowspace build-and-search hot path.
 * TEST INFRASTRUCTURE ONLY (see oracle.h for the scope statement, the reference
 * file:line each function follows and the "parity unpinned" list).
 *
 * Build:  make -C oracle      (gcc -O3 -ffp-contract=off -fopenmp)
 *
 * Every reduction below is a left-to-right loop over the contracted index, the
 * product rounded before the add.  OpenMP is only used on loops whose iterations
 * own disjoint outputs, so the result does not depend on the thread count.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TAU_FLOOR 1e-9   /* SURVEY.md Appendix A8 */

struct orc_space {
    int64_t n;
    int32_t f;
    double *items;    /* n x f, copied: helpers.rs:45 deep-copies the rows */
    double *norms;    /* n */
    double *lambdas;  /* n */
};

struct orc_graph {
    int64_t  nnodes;
    int64_t  nnz;
    int64_t *indptr;
    int32_t *indices;
    double  *data;
    orc_params   gp;
    orc_switches sw;
};

void orc_default_switches(orc_switches *sw)
{
    sw->nodes = ORC_NODES_FEATURE_COLUMNS;
    sw->kernel = ORC_KERNEL_INV_POWER;
    sw->tau_mode = ORC_TAU_MEDIAN;
    sw->lambda_form = ORC_LAMBDA_BOUNDED;
    sw->tau_fixed = 0.0;
    sw->symmetrise = ORC_SYM_MAX;
    sw->laplacian = ORC_LAPLACIAN_COMBINATORIAL;
    sw->k_counts_self = 0;
    sw->topk_prunes = 0;
    sw->distance = ORC_DISTANCE_COSINE;
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
}

/* ---------------------------------------------------------------- dot products */

/* acc[(a-a0)*m + b] = sum_t nt[t*m+a] * nt[t*m+b] for a in [a0,a1), b in [a,m): t ascending.
 * Entries b < a are left at zero (callers mirror them: the product commutes and the
 * order in t is the same, so G[b][a] is bit-identical to G[a][b]). */
static void dots_block_upper(const double *nt, int64_t d, int64_t m, int64_t a0, int64_t a1,
                             double *acc)
{
    memset(acc, 0, (size_t)((a1 - a0) * m) * sizeof(double));
    for (int64_t t = 0; t < d; ++t) {
        const double *row = nt + t * m;
        for (int64_t a = a0; a < a1; ++a) {
            const double va = row[a];
            double *A = acc + (a - a0) * m;
            for (int64_t b = a; b < m; ++b) A[b] += va * row[b];
        }
    }
}

/* Full rows (all b) for a in [a0,a1). */
static void dots_block_full(const double *nt, int64_t d, int64_t m, int64_t a0, int64_t a1,
                            double *acc)
{
    memset(acc, 0, (size_t)((a1 - a0) * m) * sizeof(double));
    for (int64_t t = 0; t < d; ++t) {
        const double *row = nt + t * m;
        for (int64_t a = a0; a < a1; ++a) {
            const double va = row[a];
            double *A = acc + (a - a0) * m;
            for (int64_t b = 0; b < m; ++b) A[b] += va * row[b];
        }
    }
}

void orc_gram_columns(const double *x, int64_t n, int64_t m, double *out)
{
    const int64_t blk = 8;
    const int64_t nblk = (m + blk - 1) / blk;
#pragma omp parallel
    {
        double *acc = (double *)malloc((size_t)(blk * m) * sizeof(double));
#pragma omp for schedule(dynamic, 1)
        for (int64_t ib = 0; ib < nblk; ++ib) {
            const int64_t a0 = ib * blk, a1 = (a0 + blk < m) ? a0 + blk : m;
            dots_block_upper(x, n, m, a0, a1, acc);
            for (int64_t a = a0; a < a1; ++a)
                for (int64_t b = a; b < m; ++b) out[a * m + b] = acc[(a - a0) * m + b];
        }
        free(acc);
    }
    for (int64_t a = 0; a < m; ++a)
        for (int64_t b = 0; b < a; ++b) out[a * m + b] = out[b * m + a];
}

static double seq_dot(const double *a, const double *b, int64_t n)
{
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

/* ---------------------------------------------------------------- neighbour selection */

typedef struct { double d; int64_t b; } cand_t;

static inline int cand_less(const cand_t *x, const cand_t *y)
{
    return (x->d < y->d) || (x->d == y->d && x->b < y->b);
}

/* max-heap on (d, b): the root is the worst kept candidate */
static void heap_sift_down(cand_t *h, int64_t n, int64_t i)
{
    for (;;) {
        int64_t l = 2 * i + 1, r = l + 1, w = i;
        if (l < n && cand_less(&h[w], &h[l])) w = l;
        if (r < n && cand_less(&h[w], &h[r])) w = r;
        if (w == i) return;
        cand_t t = h[i]; h[i] = h[w]; h[w] = t;
        i = w;
    }
}

static void heap_sift_up(cand_t *h, int64_t i)
{
    while (i > 0) {
        int64_t p = (i - 1) / 2;
        if (!cand_less(&h[p], &h[i])) return;
        cand_t t = h[i]; h[i] = h[p]; h[p] = t;
        i = p;
    }
}

static int cand_cmp(const void *x, const void *y)
{
    const cand_t *a = (const cand_t *)x, *b = (const cand_t *)y;
    if (cand_less(a, b)) return -1;
    if (cand_less(b, a)) return 1;
    return 0;
}

/* A3: distance of two nodes from their left-to-right sums <a,b>, <a,a>, <b,b>.
 * cosine (default): 1 - max(0, <a,b> / (sqrt<a,a> sqrt<b,b>)), GRAPH_VARIABLES.md:7,37.
 * l2 / l2sq [UNPINNED variants]: the Gram form (<a,a> + <b,b>) - 2<a,b>, clamped at 0, and its square root. */
static inline double node_distance(double dot, double saa, double sbb, double na, double nb, int distance)
{
    if (distance != ORC_DISTANCE_COSINE) {
        double d2 = (saa + sbb) - 2.0 * dot;
        if (!(d2 > 0.0)) d2 = 0.0;
        return distance == ORC_DISTANCE_L2 ? sqrt(d2) : d2;
    }
    double c = 0.0;                                       /* A3: c := 0 if a norm is 0 */
    if (na != 0.0 && nb != 0.0) c = dot / (na * nb);
    return 1.0 - (c > 0.0 ? c : 0.0);
}

/* A3 + A4 for one node a given its dot products with every node.
 * Writes up to kk (d, b) pairs, ascending by (d, b); returns the count. */
static int64_t select_neighbours(const double *dots, const double *norm, const double *nsq, int64_t m, int64_t a,
                                 double eps, int64_t kk, int distance, cand_t *heap)
{
    int64_t cnt = 0;
    if (kk <= 0) return 0;
    const double na = norm[a];
    for (int64_t b = 0; b < m; ++b) {
        if (b == a) continue;
        const double dist = node_distance(dots[b], nsq[a], nsq[b], na, norm[b], distance);
        if (!(dist <= eps)) continue;                     /* GRAPH_VARIABLES.md:7 */
        cand_t cd = { dist, b };
        if (cnt < kk) {                                   /* GRAPH_VARIABLES.md:8: k-cap */
            heap[cnt] = cd;
            heap_sift_up(heap, cnt);
            ++cnt;
        } else if (cand_less(&cd, &heap[0])) {
            heap[0] = cd;
            heap_sift_down(heap, cnt, 0);
        }
    }
    qsort(heap, (size_t)cnt, sizeof(cand_t), cand_cmp);
    return cnt;
}

static inline double edge_weight(double dist, const orc_params *gp, int kernel)
{
    const double t = pow(dist / gp->sigma, gp->p);
    if (kernel == ORC_KERNEL_GAUSSIAN) return exp(-t);
    return 1.0 / (1.0 + t);                               /* GRAPH_VARIABLES.md:3,9 */
}

/* ---------------------------------------------------------------- CSR assembly */

typedef struct { int64_t r; int64_t c; double w; } edge_t;

static int edge_cmp(const void *x, const void *y)
{
    const edge_t *a = (const edge_t *)x, *b = (const edge_t *)y;
    if (a->r != b->r) return a->r < b->r ? -1 : 1;
    if (a->c != b->c) return a->c < b->c ? -1 : 1;
    if (a->w != b->w) return a->w > b->w ? -1 : 1;       /* larger weight first: max rule */
    return 0;
}

static int list_has(const int64_t *nbr, const int64_t *cnt, int64_t kk, int64_t row, int64_t v)
{
    for (int64_t j = 0; j < cnt[row]; ++j)
        if (nbr[row * kk + j] == v) return 1;
    return 0;
}

/* A6 + A7: symmetrise (default W = max(W, W^T)), L = D - W (or a normalised variant), CSR with sorted columns and the
 * diagonal.  d is symmetric, so an edge selected by both endpoints carries the same weight in both lists: the rules only
 * differ on ONE-SIDED edges -- max keeps them in both directions, avg halves them, min drops them, none keeps the
 * selecting direction only. */
static int assemble_laplacian(int64_t m, int64_t kk, const int64_t *cnt, const int64_t *nbr,
                              const double *wgt, const orc_switches *sw, orc_graph *g)
{
    int64_t ne = 0;
    for (int64_t a = 0; a < m; ++a) ne += cnt[a];
    edge_t *e = (edge_t *)malloc((size_t)(2 * ne + 1) * sizeof(edge_t));
    if (!e) return ORC_ERR_NOMEM;
    int64_t q = 0;
    for (int64_t a = 0; a < m; ++a)
        for (int64_t j = 0; j < cnt[a]; ++j) {
            const int64_t b = nbr[a * kk + j];
            double w = wgt[a * kk + j];
            if (!(w > 0.0)) continue;                     /* W entries that underflowed to 0 are not edges (A7) */
            int mirror = 1;
            if (sw->symmetrise != ORC_SYM_MAX) {
                const int mutual = list_has(nbr, cnt, kk, b, a);
                if (sw->symmetrise == ORC_SYM_NONE) mirror = 0;
                else if (!mutual && sw->symmetrise == ORC_SYM_MIN) continue;
                else if (!mutual && sw->symmetrise == ORC_SYM_AVG) w = 0.5 * w;
            }
            e[q].r = a; e[q].c = b; e[q].w = w; ++q;
            if (mirror) { e[q].r = b; e[q].c = a; e[q].w = w; ++q; }
        }
    qsort(e, (size_t)q, sizeof(edge_t), edge_cmp);
    int64_t u = 0;                                        /* dedupe, first (= max) wins */
    for (int64_t i = 0; i < q; ++i)
        if (u == 0 || e[i].r != e[u - 1].r || e[i].c != e[u - 1].c) e[u++] = e[i];

    g->nnodes = m;
    g->nnz = u + m;
    g->indptr = (int64_t *)malloc((size_t)(m + 1) * sizeof(int64_t));
    g->indices = (int32_t *)malloc((size_t)(g->nnz) * sizeof(int32_t));
    g->data = (double *)malloc((size_t)(g->nnz) * sizeof(double));
    double *degs = (double *)malloc((size_t)m * sizeof(double));
    if (!g->indptr || !g->indices || !g->data || !degs) { free(e); free(degs); return ORC_ERR_NOMEM; }
    int64_t pos = 0, i = 0;
    for (int64_t a = 0; a < m; ++a) {
        g->indptr[a] = pos;
        int64_t j = i;
        double deg = 0.0;
        while (j < u && e[j].r == a) { deg += e[j].w; ++j; }   /* ascending column order */
        degs[a] = deg;
        int diag_done = 0;
        for (int64_t t = i; t < j; ++t) {
            if (!diag_done && e[t].c > a) {
                g->indices[pos] = (int32_t)a; g->data[pos] = deg; ++pos; diag_done = 1;
            }
            g->indices[pos] = (int32_t)e[t].c; g->data[pos] = -e[t].w; ++pos;
        }
        if (!diag_done) { g->indices[pos] = (int32_t)a; g->data[pos] = deg; ++pos; }
        i = j;
    }
    g->indptr[m] = pos;
    free(e);
    /* A7 variants [UNPINNED]: sym  L_ab = -w_ab / sqrt(deg_a deg_b), L_aa = [deg_a > 0];  rw  L_ab = -w_ab / deg_a.
     * Entries towards a node of degree 0 (possible with symmetrise = none) become explicit zeros. */
    if (sw->laplacian != ORC_LAPLACIAN_COMBINATORIAL)
        for (int64_t a = 0; a < m; ++a)
            for (int64_t t = g->indptr[a]; t < g->indptr[a + 1]; ++t) {
                const int64_t b = g->indices[t];
                if (b == a) { g->data[t] = degs[a] > 0.0 ? 1.0 : 0.0; continue; }
                const double w = -g->data[t];
                double v = 0.0;
                if (sw->laplacian == ORC_LAPLACIAN_SYM) { if (degs[a] > 0.0 && degs[b] > 0.0) v = -(w / sqrt(degs[a] * degs[b])); }
                else if (degs[a] > 0.0) v = -(w / degs[a]);
                g->data[t] = v;
            }
    free(degs);
    return ORC_OK;
}

int orc_graph_from_nodes_t(const double *nt, int64_t d, int64_t m, const orc_params *gp,
                           const orc_switches *sw_in, orc_graph **out_graph)
{
    if (!nt || !gp || !out_graph || d <= 0 || m <= 0) return ORC_ERR_ARG;
    if (m > 2147483647LL) return ORC_ERR_ARG;
    orc_switches sw;
    if (sw_in) sw = *sw_in; else orc_default_switches(&sw);

    /* A4 neighbour cap + the unpinned conventions: min(k, topk) when topk prunes, minus the node itself when k counts it */
    int64_t kk = gp->k;
    if (sw.topk_prunes && gp->topk < kk) kk = gp->topk;
    if (sw.k_counts_self) kk -= 1;
    if (kk > m - 1) kk = m - 1;
    if (kk < 0) kk = 0;
    const int64_t kalloc = kk > 0 ? kk : 1;

    double *norm = (double *)malloc((size_t)m * sizeof(double));
    double *nsq = (double *)malloc((size_t)m * sizeof(double));
    int64_t *cnt = (int64_t *)calloc((size_t)m, sizeof(int64_t));
    int64_t *nbr = (int64_t *)malloc((size_t)(m * kalloc) * sizeof(int64_t));
    double *wgt = (double *)malloc((size_t)(m * kalloc) * sizeof(double));
    if (!norm || !nsq || !cnt || !nbr || !wgt) { free(norm); free(nsq); free(cnt); free(nbr); free(wgt); return ORC_ERR_NOMEM; }

    /* A3: norms, sqrt of the sequential sum of squares */
#pragma omp parallel for schedule(static)
    for (int64_t a = 0; a < m; ++a) {
        double s = 0.0;
        for (int64_t t = 0; t < d; ++t) { const double v = nt[t * m + a]; s += v * v; }
        nsq[a] = s;
        norm[a] = sqrt(s);
    }

    const int64_t blk = 8;
    const int64_t nblk = (m + blk - 1) / blk;
#pragma omp parallel
    {
        double *acc = (double *)malloc((size_t)(blk * m) * sizeof(double));
        cand_t *heap = (cand_t *)malloc((size_t)kalloc * sizeof(cand_t));
#pragma omp for schedule(dynamic, 1)
        for (int64_t ib = 0; ib < nblk; ++ib) {
            const int64_t a0 = ib * blk, a1 = (a0 + blk < m) ? a0 + blk : m;
            dots_block_full(nt, d, m, a0, a1, acc);
            for (int64_t a = a0; a < a1; ++a) {
                const int64_t c = select_neighbours(acc + (a - a0) * m, norm, nsq, m, a, gp->eps, kk, sw.distance, heap);
                cnt[a] = c;
                for (int64_t j = 0; j < c; ++j) {
                    nbr[a * kalloc + j] = heap[j].b;
                    wgt[a * kalloc + j] = edge_weight(heap[j].d, gp, sw.kernel);   /* A5 */
                }
            }
        }
        free(acc);
        free(heap);
    }

    orc_graph *g = (orc_graph *)calloc(1, sizeof(orc_graph));
    if (!g) { free(norm); free(nsq); free(cnt); free(nbr); free(wgt); return ORC_ERR_NOMEM; }
    g->gp = *gp;
    g->sw = sw;
    int rc = assemble_laplacian(m, kalloc, cnt, nbr, wgt, &sw, g);
    free(norm); free(nsq); free(cnt); free(nbr); free(wgt);
    if (rc != ORC_OK) { orc_free_graph(g); return rc; }
    *out_graph = g;
    return ORC_OK;
}

/* ---------------------------------------------------------------- taumode lambda (A8) */

static int dbl_cmp(const void *x, const void *y)
{
    const double a = *(const double *)x, b = *(const double *)y;
    return (a < b) ? -1 : (a > b) ? 1 : 0;
}

static double median_of(const double *x, int64_t n, int use_abs, double *scratch)
{
    for (int64_t i = 0; i < n; ++i) scratch[i] = use_abs ? fabs(x[i]) : x[i];
    qsort(scratch, (size_t)n, sizeof(double), dbl_cmp);
    return (n & 1) ? scratch[n / 2] : 0.5 * (scratch[n / 2 - 1] + scratch[n / 2]);
}

static double tau_of(const double *x, int64_t n, const orc_switches *sw, double *scratch)
{
    double t;
    switch (sw->tau_mode) {
    case ORC_TAU_MEDIAN:     t = median_of(x, n, 0, scratch); break;
    case ORC_TAU_MEDIAN_ABS: t = median_of(x, n, 1, scratch); break;
    case ORC_TAU_MEAN: {
        double s = 0.0;
        for (int64_t i = 0; i < n; ++i) s += x[i];
        t = s / (double)n;
        break;
    }
    default: t = sw->tau_fixed; break;
    }
    return (t > TAU_FLOOR) ? t : TAU_FLOOR;
}

/* E = x^T L x / x^T x (TAUMODE.md:18,24); lambda = E/(E+tau) (TAUMODE.md:19,25). */
static int taumode_one(const orc_graph *g, const orc_switches *sw, const double *x, double *scratch,
                       double *out_e, double *out_tau, double *out_lambda)
{
    const int64_t m = g->nnodes;
    double num = 0.0;
    for (int64_t a = 0; a < m; ++a) {
        double y = 0.0;
        for (int64_t j = g->indptr[a]; j < g->indptr[a + 1]; ++j) y += g->data[j] * x[g->indices[j]];
        num += x[a] * y;
    }
    const double den = seq_dot(x, x, m);
    if (den == 0.0) return ORC_ERR_ZERO_VECTOR;           /* TAUMODE.md:13 */
    const double e = num / den;
    const double tau = tau_of(x, m, sw, scratch);
    const double eb = e / (e + tau);
    double lam = eb;
    if (sw->lambda_form == ORC_LAMBDA_SYNTHETIC) {        /* TAUMODE.md:8,26-27 */
        /* edgewise Dirichlet energies on the coefficients of the symmetrised form, c_ab = -(L_ab + L_ba)/2, for which
         * x^T L x = sum_{a<b} c_ab (x_a - x_b)^2 holds (c_ab = w_ab for the default symmetric Laplacian) */
        double tot = 0.0, sq = 0.0;
        for (int64_t a = 0; a < m; ++a)
            for (int64_t j = g->indptr[a]; j < g->indptr[a + 1]; ++j) {
                const int64_t b = g->indices[j];
                if (b == a) continue;
                double lba = 0.0;
                int have_ba = 0;
                for (int64_t t = g->indptr[b]; t < g->indptr[b + 1]; ++t)
                    if (g->indices[t] == a) { lba = g->data[t]; have_ba = 1; break; }
                if (b < a && have_ba) continue;           /* the pair was handled from row b */
                const double cab = (b > a) ? (-0.5 * g->data[j]) + (-0.5 * lba) : (-0.5 * lba) + (-0.5 * g->data[j]);
                const double df = x[a] - x[b];
                const double en = cab * (df * df);
                tot += en;
                sq += en * en;
            }
        double gd = (tot == 0.0) ? 0.0 : sq / (tot * tot);
        if (gd < 0.0) gd = 0.0;
        if (gd > 1.0) gd = 1.0;
        lam = tau * eb + (1.0 - tau) * gd;
    }
    if (out_e) *out_e = e;
    if (out_tau) *out_tau = tau;
    if (out_lambda) *out_lambda = lam;
    return ORC_OK;
}

int orc_taumode(const orc_graph *g, const orc_switches *sw_in, const double *x, int64_t nq,
                double *out_energy, double *out_tau, double *out_lambda)
{
    if (!g || !x || nq < 0) return ORC_ERR_ARG;
    orc_switches sw = sw_in ? *sw_in : g->sw;
    const int64_t m = g->nnodes;
    int rc = ORC_OK;
#pragma omp parallel
    {
        double *scratch = (double *)malloc((size_t)m * sizeof(double));
#pragma omp for schedule(static)
        for (int64_t i = 0; i < nq; ++i) {
            double e = NAN, t = NAN, l = NAN;
            const int r = taumode_one(g, &sw, x + i * m, scratch, &e, &t, &l);
            if (r != ORC_OK) {
#pragma omp atomic write
                rc = r;
            }
            if (out_energy) out_energy[i] = e;
            if (out_tau) out_tau[i] = t;
            if (out_lambda) out_lambda[i] = l;
        }
        free(scratch);
    }
    return rc;
}

/* ---------------------------------------------------------------- build (A1-A8) */

int orc_build(const double *items, int64_t n, int32_t f, const orc_params *gp,
              const orc_switches *sw_in, orc_space **out_space, orc_graph **out_graph)
{
    if (!items || !gp || !out_space || !out_graph) return ORC_ERR_ARG;
    if (n <= 0 || f <= 0) return ORC_ERR_EMPTY;           /* helpers.rs:27-29 */
    orc_switches sw;
    if (sw_in) sw = *sw_in; else orc_default_switches(&sw);

    orc_space *s = (orc_space *)calloc(1, sizeof(orc_space));
    if (!s) return ORC_ERR_NOMEM;
    s->n = n; s->f = f;
    s->items = (double *)malloc((size_t)(n * f) * sizeof(double));
    s->norms = (double *)malloc((size_t)n * sizeof(double));
    s->lambdas = (double *)malloc((size_t)n * sizeof(double));
    if (!s->items || !s->norms || !s->lambdas) { orc_free_space(s); return ORC_ERR_NOMEM; }
    memcpy(s->items, items, (size_t)(n * f) * sizeof(double));   /* A1: stored unchanged */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i)
        s->norms[i] = sqrt(seq_dot(s->items + i * f, s->items + i * f, f));

    orc_graph *g = NULL;
    int rc;
    if (sw.nodes == ORC_NODES_FEATURE_COLUMNS) {
        /* A2: node a = column a of X; the transposed node matrix IS X (d = n, m = f). */
        rc = orc_graph_from_nodes_t(s->items, n, f, gp, &sw, &g);
        if (rc == ORC_OK) rc = orc_taumode(g, &sw, s->items, n, NULL, NULL, s->lambdas);
    } else {
        double *xt = (double *)malloc((size_t)(n * f) * sizeof(double));
        if (!xt) { orc_free_space(s); return ORC_ERR_NOMEM; }
        for (int64_t i = 0; i < n; ++i)
            for (int64_t t = 0; t < f; ++t) xt[t * n + i] = s->items[i * f + t];
        rc = orc_graph_from_nodes_t(xt, f, n, gp, &sw, &g);
        free(xt);
        for (int64_t i = 0; i < n; ++i) s->lambdas[i] = NAN;
    }
    if (rc != ORC_OK) { orc_free_space(s); orc_free_graph(g); return rc; }
    *out_space = s;
    *out_graph = g;
    return ORC_OK;
}

/* ---------------------------------------------------------------- pre-graph reduction (R1-R6) */

void orc_default_reduction(orc_reduction *red)
{
    red->sample_rate = 0.6;   /* suggested_eps.md:6 "keep rate 60.0%" */
    red->seed = 42;           /* src/lib.rs:283 */
    red->n_clusters = 0;
    red->max_iters = 10;
    red->probes = 2048;
    red->reserved = 0;
}

static uint64_t splitmix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

int64_t orc_reduction_sample(const orc_reduction *red, int64_t row0, int64_t n, int32_t *out_rows)
{
    int64_t cnt = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (red->sample_rate < 1.0) {
            const uint64_t z = splitmix64(red->seed + (uint64_t)(row0 + i + 1) * 0x9E3779B97F4A7C15ULL);
            const double u = (double)(z >> 11) * (1.0 / 9007199254740992.0);
            if (!(u < red->sample_rate)) continue;
        }
        out_rows[cnt++] = (int32_t)i;
    }
    return cnt;
}

/* squared Euclidean distance, left to right: difference, product rounded, then added */
static double seq_sqdist(const double *a, const double *b, int64_t f)
{
    double s = 0.0;
    for (int64_t t = 0; t < f; ++t) {
        const double d = a[t] - b[t];
        s += d * d;
    }
    return s;
}

int orc_reduce(const double *items, int64_t n, int32_t f, const orc_reduction *red_in, int64_t n_total_for_k,
               orc_reduction_info *info, double *centroids_out, int64_t cap_clusters)
{
    if (!items || !centroids_out) return ORC_ERR_ARG;
    if (n <= 0 || f <= 0) return ORC_ERR_EMPTY;
    orc_reduction red;
    if (red_in) red = *red_in; else orc_default_reduction(&red);
    if (!(red.sample_rate > 0.0) || red.max_iters < 0 || red.n_clusters < 0 || red.probes < 0) return ORC_ERR_ARG;

    /* R1 */
    int32_t *rows = (int32_t *)malloc((size_t)n * sizeof(int32_t));
    if (!rows) return ORC_ERR_NOMEM;
    int64_t ns = orc_reduction_sample(&red, 0, n, rows);
    if (ns == 0) { for (int64_t i = 0; i < n; ++i) rows[i] = (int32_t)i; ns = n; }   /* an empty sample keeps every row */

    orc_reduction_info inf;
    memset(&inf, 0, sizeof(inf));
    inf.n_sampled = ns;
    inf.two_nn_mean_ratio = NAN;

    /* R2 */
    if (red.probes > 0 && ns >= 3) {
        const int64_t P = red.probes < ns ? red.probes : ns;
        double *r12 = (double *)malloc((size_t)(2 * P) * sizeof(double));
        if (!r12) { free(rows); return ORC_ERR_NOMEM; }
#pragma omp parallel for schedule(dynamic, 4)
        for (int64_t j = 0; j < P; ++j) {
            const int64_t pos = (j * ns) / P;
            const double *x = items + (int64_t)rows[pos] * f;
            double d1 = INFINITY, d2 = INFINITY;              /* ties by position: a later equal distance never displaces */
            for (int64_t b = 0; b < ns; ++b) {
                if (b == pos) continue;
                const double d = seq_sqdist(x, items + (int64_t)rows[b] * f, f);
                if (d < d1) { d2 = d1; d1 = d; }
                else if (d < d2) d2 = d;
            }
            r12[2 * j] = d1; r12[2 * j + 1] = d2;
        }
        double acc = 0.0;
        int64_t used = 0;
        for (int64_t j = 0; j < P; ++j) {
            const double r1 = sqrt(r12[2 * j]), r2 = sqrt(r12[2 * j + 1]);
            if (!(r1 > 0.0) || !isfinite(r2)) continue;        /* duplicates carry no scale information */
            acc += r2 / r1;
            ++used;
        }
        free(r12);
        inf.n_probes = used;
        if (used > 0) {
            const double m = acc / (double)used;
            inf.two_nn_mean_ratio = m;
            double d = (m > 1.0) ? m / (m - 1.0) : (double)f;  /* mean of a Pareto(d) ratio is d/(d-1); 1.3560 -> 3 (suggested_eps.md:9) */
            if (!(d < (double)f)) d = (double)f;
            inf.intrinsic_dim = d < 1.0 ? 1 : (int32_t)d;
        }
    }

    /* R3: the one published data point is N = 313841 -> "Testing K in range [178, 179]" (suggested_eps.md:11) */
    int64_t K = red.n_clusters;
    if (K <= 0) {
        const int64_t N = n_total_for_k > 0 ? n_total_for_k : n;
        K = (int64_t)ceil(sqrt((double)N / 10.0));
    }
    if (K > ns) K = ns;
    if (K < 1) K = 1;
    if (K > cap_clusters) { free(rows); return ORC_ERR_ARG; }
    inf.n_clusters = (int32_t)K;

    /* R4 */
    double *C = centroids_out;
    for (int64_t j = 0; j < K; ++j)
        memcpy(C + j * f, items + (int64_t)rows[(j * ns) / K] * f, (size_t)f * sizeof(double));
    int32_t *assign = (int32_t *)malloc((size_t)ns * sizeof(int32_t));
    double *sum = (double *)malloc((size_t)f * sizeof(double));
    if (!assign || !sum) { free(rows); free(assign); free(sum); return ORC_ERR_NOMEM; }
    for (int64_t i = 0; i < ns; ++i) assign[i] = -1;
    for (int it = 0; it < red.max_iters; ++it) {
        int64_t changed = 0;
#pragma omp parallel for schedule(static) reduction(+ : changed)
        for (int64_t i = 0; i < ns; ++i) {
            const double *x = items + (int64_t)rows[i] * f;
            double best = INFINITY;
            int32_t bj = 0;
            for (int64_t j = 0; j < K; ++j) {
                const double d = seq_sqdist(x, C + j * f, f);
                if (d < best) { best = d; bj = (int32_t)j; }   /* ties -> smaller centroid */
            }
            if (bj != assign[i]) { assign[i] = bj; ++changed; }
        }
        if (changed == 0) { inf.converged = 1; break; }
        for (int64_t j = 0; j < K; ++j) {                      /* rows added in ascending order, then one division */
            int64_t cnt = 0;
            for (int32_t t = 0; t < f; ++t) sum[t] = 0.0;
            for (int64_t i = 0; i < ns; ++i) {
                if (assign[i] != (int32_t)j) continue;
                const double *x = items + (int64_t)rows[i] * f;
                for (int32_t t = 0; t < f; ++t) sum[t] += x[t];
                ++cnt;
            }
            if (cnt > 0)
                for (int32_t t = 0; t < f; ++t) C[j * f + t] = sum[t] / (double)cnt;
        }
        inf.iters = it + 1;
    }
    free(rows); free(assign); free(sum);
    if (info) *info = inf;
    return ORC_OK;
}

int orc_build_reduced(const double *items, int64_t n, int32_t f, const orc_params *gp, const orc_switches *sw_in,
                      const orc_reduction *red, orc_space **out_space, orc_graph **out_graph, orc_reduction_info *info,
                      double *centroids_out, int64_t cap_clusters)
{
    if (!items || !gp || !out_space || !out_graph) return ORC_ERR_ARG;
    if (n <= 0 || f <= 0) return ORC_ERR_EMPTY;
    orc_switches sw;
    if (sw_in) sw = *sw_in; else orc_default_switches(&sw);
    if (sw.nodes != ORC_NODES_FEATURE_COLUMNS) return ORC_ERR_ARG;
    orc_reduction_info inf;
    int64_t cap = cap_clusters;
    double *C = centroids_out;
    if (!C) {
        cap = n < 65535 ? n : 65535;
        C = (double *)malloc((size_t)(cap * f) * sizeof(double));
        if (!C) return ORC_ERR_NOMEM;
    }
    int rc = orc_reduce(items, n, f, red, n, &inf, C, cap);
    orc_space *s = NULL;
    orc_graph *g = NULL;
    if (rc == ORC_OK) {
        s = (orc_space *)calloc(1, sizeof(orc_space));
        if (!s) rc = ORC_ERR_NOMEM;
    }
    if (rc == ORC_OK) {
        s->n = n; s->f = f;
        s->items = (double *)malloc((size_t)(n * f) * sizeof(double));
        s->norms = (double *)malloc((size_t)n * sizeof(double));
        s->lambdas = (double *)malloc((size_t)n * sizeof(double));
        if (!s->items || !s->norms || !s->lambdas) rc = ORC_ERR_NOMEM;
    }
    if (rc == ORC_OK) {
        memcpy(s->items, items, (size_t)(n * f) * sizeof(double));
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i)
            s->norms[i] = sqrt(seq_dot(s->items + i * f, s->items + i * f, f));
        /* R5: node a = column a of the centroid matrix (K x f row-major IS the transposed node matrix, d = K, m = f) */
        rc = orc_graph_from_nodes_t(C, inf.n_clusters, f, gp, &sw, &g);
    }
    if (rc == ORC_OK) rc = orc_taumode(g, &sw, s->items, n, NULL, NULL, s->lambdas);   /* R6 */
    if (!centroids_out) free(C);
    if (rc != ORC_OK) { orc_free_space(s); orc_free_graph(g); return rc; }
    if (info) *info = inf;
    *out_space = s;
    *out_graph = g;
    return ORC_OK;
}

/* ---------------------------------------------------------------- search (A9) */

void orc_scores(const orc_space *s, const double *q, double lambda_q, double tau, double *out)
{
    const int64_t n = s->n, f = s->f;
    const double nq = sqrt(seq_dot(q, q, f));
    for (int64_t i = 0; i < n; ++i) {
        const double den = nq * s->norms[i];
        const double c = (den == 0.0) ? 0.0 : seq_dot(q, s->items + i * f, f) / den;   /* README KAT */
        out[i] = tau * c + (1.0 - tau) * (1.0 / (1.0 + fabs(lambda_q - s->lambdas[i]))); /* TAUMODE.md:33 */
    }
}

typedef struct { double s; int64_t i; } hit_t;

/* "better" = larger score, ties -> smaller index */
static inline int hit_better(const hit_t *x, const hit_t *y)
{
    return (x->s > y->s) || (x->s == y->s && x->i < y->i);
}

static int hit_cmp(const void *x, const void *y)
{
    const hit_t *a = (const hit_t *)x, *b = (const hit_t *)y;
    if (hit_better(a, b)) return -1;
    if (hit_better(b, a)) return 1;
    return 0;
}

int orc_search(const orc_space *s, const orc_graph *g, const orc_switches *sw_in, const double *q,
               int64_t nq, double tau, int64_t *out_idx, double *out_score, double *out_lambda_q)
{
    if (!s || !g || !q || nq < 0 || !out_idx || !out_score) return ORC_ERR_ARG;
    if (g->nnodes != s->f) return ORC_ERR_ARG;
    orc_switches sw = sw_in ? *sw_in : g->sw;
    const int64_t n = s->n, f = s->f;
    const int64_t topk = g->gp.topk;                      /* lib.rs:169 */
    const int64_t kk = topk < n ? topk : n;
    int rc = ORC_OK;
#pragma omp parallel
    {
        double *scratch = (double *)malloc((size_t)f * sizeof(double));
        double *sc = (double *)malloc((size_t)n * sizeof(double));
        hit_t *heap = (hit_t *)malloc((size_t)(kk > 0 ? kk : 1) * sizeof(hit_t));
#pragma omp for schedule(dynamic, 1)
        for (int64_t qi = 0; qi < nq; ++qi) {
            const double *qv = q + qi * f;
            for (int64_t j = 0; j < topk; ++j) { out_idx[qi * topk + j] = -1; out_score[qi * topk + j] = NAN; }
            double lq = NAN;
            int r = taumode_one(g, &sw, qv, scratch, NULL, NULL, &lq);   /* lib.rs:154 */
            if (out_lambda_q) out_lambda_q[qi] = lq;
            if (r == ORC_OK && lq == 0.0) r = ORC_ERR_LAMBDA_ZERO;       /* lib.rs:156-159 */
            if (r != ORC_OK) {
#pragma omp atomic write
                rc = r;
                continue;
            }
            orc_scores(s, qv, lq, tau, sc);
            /* min-heap of the kk best: root = worst kept */
            int64_t cnt = 0;
            for (int64_t i = 0; i < n && kk > 0; ++i) {
                hit_t h = { sc[i], i };
                if (cnt < kk) {
                    int64_t c = cnt++;
                    heap[c] = h;
                    while (c > 0) {
                        int64_t p = (c - 1) / 2;
                        if (!hit_better(&heap[p], &heap[c])) break;
                        hit_t t = heap[p]; heap[p] = heap[c]; heap[c] = t; c = p;
                    }
                } else if (hit_better(&h, &heap[0])) {
                    heap[0] = h;
                    int64_t c = 0;
                    for (;;) {
                        int64_t l = 2 * c + 1, rr = l + 1, w = c;
                        if (l < cnt && hit_better(&heap[w], &heap[l])) w = l;
                        if (rr < cnt && hit_better(&heap[w], &heap[rr])) w = rr;
                        if (w == c) break;
                        hit_t t = heap[w]; heap[w] = heap[c]; heap[c] = t; c = w;
                    }
                }
            }
            qsort(heap, (size_t)cnt, sizeof(hit_t), hit_cmp);
            for (int64_t j = 0; j < cnt; ++j) { out_idx[qi * topk + j] = heap[j].i; out_score[qi * topk + j] = heap[j].s; }
        }
        free(scratch); free(sc); free(heap);
    }
    return rc;
}

/* ---------------------------------------------------------------- hybrid search
 *
 * ArrowSpace.search_hybrid (src/lib.rs:182-219): lambda_q = prepare_query_item (lib.rs:205), k = gl.graph_params.topk
 * (lib.rs:214), then the crate's search_lambda_aware_hybrid(&query, k, tau) (lib.rs:218).  That function's body is in the
 * un-vendored crate and nothing under /root/reference documents or tests it: PARITY UNPINNED.  Restated as the two-stage
 * reading of "hybrid" -- a cosine shortlist re-ranked by the lambda-aware score -- with the shortlist length an explicit
 * argument:
 *   H1  lambda_q as in search; NO lambda_q != 0 assertion (search_hybrid has none, unlike lib.rs:156-159);
 *   H2  shortlist = the `pool` items of largest cosine (README arithmetic), ties -> smaller index; pool <= 0 means min(2 * topk, 31),
 *       and pool is raised to topk and cut to n;
 *   H3  score_i = tau * cos_i + (1 - tau) / (1 + |lambda_q - lambda_i|) (TAUMODE.md:33) over the shortlist; the best
 *       min(topk, n) by (score desc, index asc).
 * With pool >= n it is search without the assertion. */
int orc_search_hybrid(const orc_space *s, const orc_graph *g, const orc_switches *sw_in, const double *q, int64_t nq,
                      double tau, int64_t pool, int64_t *out_idx, double *out_score, double *out_lambda_q)
{
    if (!s || !g || !q || nq < 0 || !out_idx || !out_score) return ORC_ERR_ARG;
    if (g->nnodes != s->f) return ORC_ERR_ARG;
    orc_switches sw = sw_in ? *sw_in : g->sw;
    const int64_t n = s->n, f = s->f;
    const int64_t topk = g->gp.topk;                      /* lib.rs:214 */
    const int64_t kk = topk < n ? topk : n;
    int64_t m = pool > 0 ? pool : (2 * topk < 31 ? 2 * topk : 31);   /* 31: the longest list the GPU's tensor-core pass keeps */
    if (m < topk) m = topk;
    if (m > n) m = n;
    int rc = ORC_OK;
#pragma omp parallel
    {
        double *scratch = (double *)malloc((size_t)f * sizeof(double));
        hit_t *all = (hit_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(hit_t));
#pragma omp for schedule(dynamic, 1)
        for (int64_t qi = 0; qi < nq; ++qi) {
            const double *qv = q + qi * f;
            for (int64_t j = 0; j < topk; ++j) { out_idx[qi * topk + j] = -1; out_score[qi * topk + j] = NAN; }
            double lq = NAN;
            int r = taumode_one(g, &sw, qv, scratch, NULL, NULL, &lq);   /* lib.rs:205 */
            if (out_lambda_q) out_lambda_q[qi] = lq;
            if (r != ORC_OK) {
#pragma omp atomic write
                rc = r;
                continue;
            }
            const double nrm = sqrt(seq_dot(qv, qv, f));
            for (int64_t i = 0; i < n; ++i) {                            /* H2: cosine of every item */
                const double den = nrm * s->norms[i];
                all[i].s = (den == 0.0) ? 0.0 : seq_dot(qv, s->items + i * f, f) / den;
                all[i].i = i;
            }
            qsort(all, (size_t)n, sizeof(hit_t), hit_cmp);
            for (int64_t j = 0; j < m; ++j) {                            /* H3: re-rank the shortlist */
                const int64_t i = all[j].i;
                all[j].s = tau * all[j].s + (1.0 - tau) * (1.0 / (1.0 + fabs(lq - s->lambdas[i])));
            }
            qsort(all, (size_t)m, sizeof(hit_t), hit_cmp);
            for (int64_t j = 0; j < kk; ++j) { out_idx[qi * topk + j] = all[j].i; out_score[qi * topk + j] = all[j].s; }
        }
        free(scratch); free(all);
    }
    return rc;
}

/* ---------------------------------------------------------------- accessors */

int64_t orc_space_nitems(const orc_space *s) { return s->n; }
int32_t orc_space_nfeatures(const orc_space *s) { return s->f; }
const double *orc_space_items(const orc_space *s) { return s->items; }
const double *orc_space_lambdas(const orc_space *s) { return s->lambdas; }
const double *orc_space_norms(const orc_space *s) { return s->norms; }
int64_t orc_graph_nnodes(const orc_graph *g) { return g->nnodes; }
int64_t orc_graph_nnz(const orc_graph *g) { return g->nnz; }
const int64_t *orc_graph_indptr(const orc_graph *g) { return g->indptr; }
const int32_t *orc_graph_indices(const orc_graph *g) { return g->indices; }
const double *orc_graph_data(const orc_graph *g) { return g->data; }
void orc_graph_params(const orc_graph *g, orc_params *gp) { *gp = g->gp; }

void orc_free_space(orc_space *s)
{
    if (!s) return;
    free(s->items); free(s->norms); free(s->lambdas); free(s);
}

void orc_free_graph(orc_graph *g)
{
    if (!g) return;
    free(g->indptr); free(g->indices); free(g->data); free(g);
}
