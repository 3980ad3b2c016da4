/*
 * oracle/bigclam_oracle.c — CPU restatement of the reference's BigCLAM hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (bigclam_apachespark_b200/,
 * include/, libbigclam_b200.so) may include, link, import or call this file.  Allowed
 * users: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 *
 * PARITY UNPINNED: the reference (thangdnsf/BigCLAM-ApacheSpark @ 4cdee5f) ships no tests,
 * golden vectors or recorded outputs for this path, and cannot be executed here (no JVM /
 * Scala / Spark).  This file is a line-by-line restatement of the Scala; it is cross-checked
 * against an independent NumPy restatement (oracle/numpy_twin.py) and the self-consistency
 * properties in tests/, not against outputs of the reference itself.  Its formulas (not its
 * clamps or its rounding) are anchored on the published algorithm: tests/test_oracle_formula.py
 * compares LLH, gradient, per-node term and accepted step with the BigCLAM objective of the
 * thesis (eq. 3.6-3.11) evaluated over all node pairs.
 *
 * What is restated (all citations relative to /root/reference):
 *   codes/bigclam4-7.scala:28-34    step-size list by repeated `*= beta`          -> oracle_step_sizes
 *   codes/bigclam4-7.scala:39-43    MIN_P_/MAX_P_/MIN_F_/MAX_F_                    -> oracle_params
 *   codes/bigclam4-7.scala:110-113  step(): clamp(Fu + s*dir, MIN_F, MAX_F)        -> clamp_step
 *   codes/bigclam4-7.scala:135-145  Flookup: missing key == zero row               -> dense storage
 *   codes/bigclam4-7.scala:157-169  PRE block: grad_u, llh_u                       -> pre_node
 *   codes/bigclam4-7.scala:172-184  LS block: 16 candidates, Armijo, max passing   -> linesearch_node
 *   codes/bigclam4-7.scala:186-193  UPDATE: swap rows, sumF -= (sum old - sum new) -> oracle_step
 *   codes/bigclam4-7.scala:194-219  LLH with new F / new sumF                      -> oracle_llh
 *   codes/bigclamv3-7.scala:106-120 loglikelihood() (same arithmetic)              -> oracle_llh
 *   codes/bigclam4-7.scala:225-243  SGDFindC outer loop                            -> oracle_run (variant 4)
 *   codes/bigclamv3-7.scala:206-222 MBSGD (LLHold = 0.0)                           -> oracle_run (variant 3)
 *   codes/Bigclamv2.scala:203-219   MBSGD (LLHold = loglikelihood())               -> oracle_run (variant 2)
 *
 * Arithmetic order follows the Scala expressions literally (see comments at each site):
 * dots are plain left-to-right sums (Breeze 1xK * Kx1 -> reference dgemm), neighbour
 * reductions are left folds in CSR order (Array.reduce == reduceLeft), everything is fp64.
 *
 * Deviation, documented: a node with an empty neighbour list makes the reference throw
 * (empty.reduce at :167/:180/:218).  Here such a node is never updated and contributes
 * (-fu.sumF + fu.fu) to the LLH.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int32_t k;          /* K.value                      bigclam4-7.scala:134,249 */
    int32_t max_inter;  /* MaxInter = 15 -> 16 steps    bigclam4-7.scala:26      */
    double alpha;       /* 0.05                         bigclam4-7.scala:22      */
    double beta;        /* 0.1                          bigclam4-7.scala:24      */
    double min_p;       /* MIN_P_ = 0.0001              bigclam4-7.scala:40      */
    double max_p;       /* MAX_P_ = 0.9999              bigclam4-7.scala:41      */
    double min_f;       /* MIN_F_ = 0.0                 bigclam4-7.scala:42      */
    double max_f;       /* MAX_F_ = 1000.0              bigclam4-7.scala:43      */
} oracle_params;

/* bigclam4-7.scala:28-34 — listSearch = [1.0, 1.0*beta, (1.0*beta)*beta, ...], built by
 * repeated multiplication (NOT pow): out[0] = 1.0, out[j] = out[j-1] * beta. */
void oracle_step_sizes(double beta, int32_t max_inter, double *out) {
    double s = 1.0;
    out[0] = s;
    for (int i = 1; i <= max_inter; ++i) { s *= beta; out[i] = s; }
}

static inline double dot_seq(const double *a, const double *b, int k) {
    double acc = 0.0;
    for (int i = 0; i < k; ++i) acc += a[i] * b[i];
    return acc;
}

/* log(1 - clamp(exp(-x))) + x      bigclam4-7.scala:166-167,179,218 */
static inline double edge_term(double x, const oracle_params *p, double *one_minus_p) {
    double pr = fmin(fmax(exp(-x), p->min_p), p->max_p);
    if (one_minus_p) *one_minus_p = 1.0 - pr;
    return log(1.0 - pr) + x;
}

/* PRE block, bigclam4-7.scala:157-169.  Returns llh_u, writes grad (k). */
static double pre_node(const int64_t *rowptr, const int32_t *col, int64_t u,
                       const double *F, const double *sumF, const oracle_params *p,
                       double *grad /*k*/) {
    const int k = p->k;
    const double *fu = F + (size_t)u * k;
    double fusfT = dot_seq(fu, sumF, k);            /* :159 */
    double fufuT = dot_seq(fu, fu, k);              /* :161 */
    double s1 = 0.0;                                 /* kq._1 */
    for (int i = 0; i < k; ++i) grad[i] = 0.0;       /* kq._2 accumulator */
    int first = 1;
    for (int64_t e = rowptr[u]; e < rowptr[u + 1]; ++e) {
        const double *fv = F + (size_t)col[e] * k;
        double x = dot_seq(fu, fv, k);               /* :165 */
        double omp_;
        double t = edge_term(x, p, &omp_);           /* :166-167 */
        double w = 1.0 / omp_;                       /* (1/(1 - fufvTrange)) */
        if (first) {                                 /* reduce == reduceLeft: first element is the seed */
            s1 = t;
            for (int i = 0; i < k; ++i) grad[i] = fv[i] * w;
            first = 0;
        } else {
            s1 = s1 + t;
            for (int i = 0; i < k; ++i) grad[i] = grad[i] + fv[i] * w;
        }
    }
    /* (ux, kq._2 - sf + fu, kq._1 - fusfT + fufuT)   :168 */
    for (int i = 0; i < k; ++i) grad[i] = (grad[i] - sumF[i]) + fu[i];
    return (s1 - fusfT) + fufuT;
}

/* test hook: when set, ls_trial also stores result - (llh_u + arm) here (oracle_armijo_margins, single-threaded) */
static double *g_margin_out = NULL;
#ifdef _OPENMP
#pragma omp threadprivate(g_margin_out)
#endif

/* One candidate of the LS block, bigclam4-7.scala:173-181.  newfu (k) is written. */
static int ls_trial(const int64_t *rowptr, const int32_t *col, int64_t u,
                    const double *F, const double *sumF, const oracle_params *p,
                    const double *grad, double llh_u, double s,
                    double *newfu /*k*/, double *sfT /*k*/) {
    const int k = p->k;
    const double *fu = F + (size_t)u * k;
    for (int i = 0; i < k; ++i) {                    /* step(), :110-113 */
        double x = fu[i] + s * grad[i];
        newfu[i] = fmin(fmax(x, p->min_f), p->max_f);
    }
    for (int i = 0; i < k; ++i) sfT[i] = (sumF[i] - fu[i]) + newfu[i];   /* :176 */
    double acc = 0.0;
    int first = 1;
    for (int64_t e = rowptr[u]; e < rowptr[u + 1]; ++e) {                /* :177-180 */
        const double *fv = F + (size_t)col[e] * k;
        double xc = dot_seq(newfu, fv, k);
        double t = edge_term(xc, p, NULL);
        if (first) { acc = t; first = 0; } else acc = acc + t;
    }
    double result = (acc - dot_seq(newfu, sfT, k)) + dot_seq(newfu, newfu, k);
    /* (alpha*stepx*x._2 * BDM.create(K,1,x._2.data)).data.reduce(_+_)   :181
     * == sum_i ((alpha*s)*g_i)*g_i  (scalar*matrix first, then 1xK * Kx1). */
    double as = p->alpha * s;
    double arm = 0.0;
    for (int i = 0; i < k; ++i) arm += (as * grad[i]) * grad[i];
    if (g_margin_out) *g_margin_out = result - (llh_u + arm);
    return result >= (llh_u + arm);
}

/* Test helper (SURVEY 8c property iv): for `count` nodes, the Armijo margins  llh'(s_j) - (llh_u + alpha s_j |g|^2)
 * of all candidates (bigclam4-7.scala:181) as this restatement computes them, and llh_u.  A node whose accepted
 * index differs between two implementations is a genuine tie only if the margin at the first index where they
 * disagree is at rounding level. */
void oracle_armijo_margins(int64_t n, const int64_t *rowptr, const int32_t *col, const oracle_params *p,
                           const double *F, const double *sumF, const int64_t *nodes, int64_t count,
                           double *margins_out /* count x (max_inter+1) */, double *llh_u_out /* count */) {
    (void)n;
    const int k = p->k;
    const int nsteps = p->max_inter + 1;
    double *steps = (double *)malloc(sizeof(double) * (size_t)nsteps);
    oracle_step_sizes(p->beta, p->max_inter, steps);
    double *grad = (double *)malloc(sizeof(double) * (size_t)k * 3);
    double *newfu = grad + k, *sfT = grad + 2 * (size_t)k;
    for (int64_t i = 0; i < count; ++i) {
        const int64_t u = nodes[i];
        const double llh_u = pre_node(rowptr, col, u, F, sumF, p, grad);
        llh_u_out[i] = llh_u;
        for (int j = 0; j < nsteps; ++j) {
            double m = 0.0;
            g_margin_out = &m;
            (void)ls_trial(rowptr, col, u, F, sumF, p, grad, llh_u, steps[j], newfu, sfT);
            g_margin_out = NULL;
            margins_out[i * nsteps + j] = m;
        }
    }
    free(grad);
    free(steps);
}

/* per-node LLH term of the LLH block, bigclam4-7.scala:196-219 (== loglikelihood(),
 * bigclamv3-7.scala:106-120). */
static double llh_node(const int64_t *rowptr, const int32_t *col, int64_t u,
                       const double *F, const double *sumF, const oracle_params *p) {
    const int k = p->k;
    const double *fu = F + (size_t)u * k;
    double fusfT = dot_seq(fu, sumF, k);
    double fufuT = dot_seq(fu, fu, k);
    double acc = 0.0;
    int first = 1;
    for (int64_t e = rowptr[u]; e < rowptr[u + 1]; ++e) {
        const double *fv = F + (size_t)col[e] * k;
        double x = dot_seq(fu, fv, k);
        double t = edge_term(x, p, NULL);
        if (first) { acc = t; first = 0; } else acc = acc + t;
    }
    return (acc - fusfT) + fufuT;
}

/* LLH = sum_u llh_node(u), summed in node order (Spark's reduce order is unspecified). */
double oracle_llh(int64_t n, const int64_t *rowptr, const int32_t *col,
                  const oracle_params *p, const double *F, const double *sumF,
                  double *per_node /* optional n */) {
    double *tmp = per_node ? per_node : (double *)malloc(sizeof(double) * (size_t)n);
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t u = 0; u < n; ++u) tmp[u] = llh_node(rowptr, col, u, F, sumF, p);
    double llh = 0.0;
    for (int64_t u = 0; u < n; ++u) llh += tmp[u];
    if (!per_node) free(tmp);
    return llh;
}

/*
 * One call of backtrackingLineSearchs(uset)  (bigclam4-7.scala:152-223).
 *   F_in     n x k row-major, read only (the broadcast snapshot Fbc, :154 — Jacobi)
 *   sumF     k, updated in place (:192)
 *   node_mask  NULL == all nodes (uset is always all vertices, :227); else u is in uset iff mask[u] != 0
 *   F_out    n x k, receives the full new F (rows not accepted are copied, :190)
 *   accepted optional n: index j of the accepted step size (s = steps[j]) or -1
 *   trials_out optional n: number of candidates actually evaluated (16 unless early_exit)
 *   early_exit != 0: evaluate candidates in descending order and stop at the first pass
 *                    (identical result: the reference keeps the max passing step, :182)
 *   grad_out/llh_u_out optional n*k / n: the PRE block's outputs, for kernel unit tests
 * Returns the LLH after the update (:196-219).
 */
double oracle_step(int64_t n, const int64_t *rowptr, const int32_t *col,
                   const oracle_params *p, const double *F_in, double *sumF,
                   const uint8_t *node_mask, double *F_out,
                   int64_t *n_updated_out, int8_t *accepted, int8_t *trials_out,
                   int32_t early_exit, double *grad_out, double *llh_u_out) {
    const int k = p->k;
    const int nsteps = p->max_inter + 1;
    double *steps = (double *)malloc(sizeof(double) * (size_t)nsteps);
    oracle_step_sizes(p->beta, p->max_inter, steps);
    int8_t *acc_idx = accepted ? accepted : (int8_t *)malloc((size_t)n);

#pragma omp parallel
    {
        double *grad = (double *)malloc(sizeof(double) * (size_t)k * 4);
        double *newfu = grad + k, *sfT = grad + 2 * (size_t)k, *best = grad + 3 * (size_t)k;
#pragma omp for schedule(dynamic, 64)
        for (int64_t u = 0; u < n; ++u) {
            const double *fu = F_in + (size_t)u * k;
            double *out = F_out + (size_t)u * k;
            int8_t chosen = -1, ntr = 0;
            int in_uset = (node_mask == NULL) || node_mask[u];
            if (in_uset && rowptr[u + 1] > rowptr[u]) {
                double llh_u = pre_node(rowptr, col, u, F_in, sumF, p, grad);
                if (grad_out) memcpy(grad_out + (size_t)u * k, grad, sizeof(double) * (size_t)k);
                if (llh_u_out) llh_u_out[u] = llh_u;
                /* candidates in descending order: steps[0] = 1.0 is the largest;
                 * "max passing" (:182) == first passing in this order. */
                for (int j = 0; j < nsteps; ++j) {
                    ++ntr;
                    int pass = ls_trial(rowptr, col, u, F_in, sumF, p, grad, llh_u, steps[j], newfu, sfT);
                    if (pass && chosen < 0) {
                        chosen = (int8_t)j;
                        memcpy(best, newfu, sizeof(double) * (size_t)k);   /* == step(fu, s*, grad), :183 */
                        if (early_exit) break;
                    }
                }
            } else {
                if (grad_out) memset(grad_out + (size_t)u * k, 0, sizeof(double) * (size_t)k);
                if (llh_u_out) llh_u_out[u] = llh_node(rowptr, col, u, F_in, sumF, p);
            }
            acc_idx[u] = chosen;
            if (trials_out) trials_out[u] = ntr;
            memcpy(out, chosen >= 0 ? best : fu, sizeof(double) * (size_t)k);
        }
        free(grad);
    }

    /* UPDATE, :186-193: changeFu = (sum of old rows, sum of new rows) over Sx, node order;
     * sumF = sumF - (changeFu._1 - changeFu._2). */
    int64_t n_upd = 0;
    double *A = (double *)calloc((size_t)k * 2, sizeof(double));
    double *B = A + k;
    for (int64_t u = 0; u < n; ++u) {
        if (acc_idx[u] < 0) continue;
        const double *o = F_in + (size_t)u * k, *nw = F_out + (size_t)u * k;
        if (n_upd == 0) { for (int i = 0; i < k; ++i) { A[i] = o[i]; B[i] = nw[i]; } }
        else            { for (int i = 0; i < k; ++i) { A[i] = A[i] + o[i]; B[i] = B[i] + nw[i]; } }
        ++n_upd;
    }
    if (n_upd > 0) for (int i = 0; i < k; ++i) sumF[i] = sumF[i] - (A[i] - B[i]);
    free(A);
    if (n_updated_out) *n_updated_out = n_upd;
    if (!accepted) free(acc_idx);
    free(steps);

    return oracle_llh(n, rowptr, col, p, F_out, sumF, NULL);
}

/* sumF = exact column sums of F in node order (bigclam4-7.scala:105-106; Bigclamv2.scala:95). */
void oracle_colsum(int64_t n, int32_t k, const double *F, double *sumF) {
    for (int i = 0; i < k; ++i) sumF[i] = 0.0;
    for (int64_t u = 0; u < n; ++u)
        for (int i = 0; i < k; ++i) sumF[i] += F[(size_t)u * k + i];
}

/*
 * Outer loop.  variant 4: SGDFindC (bigclam4-7.scala:225-243): LLHold = one step; loop
 * {new = step; if |1 - new/old| < tol break; old = new}; returns LLHold (the value BEFORE the
 * converged one, as coded at :242).  variant 3: MBSGD with LLHold = 0.0 (bigclamv3-7.scala:207).
 * variant 2: MBSGD with LLHold = loglikelihood() (Bigclamv2.scala:204).
 * max_outer == 0 means unbounded like the reference; F is updated in place (n x k).
 * llh_trace (optional, capacity trace_cap) receives every step's returned LLH.
 * Returns the number of hot-path calls made; *llh_out as described.
 */
int64_t oracle_run(int64_t n, const int64_t *rowptr, const int32_t *col,
                   const oracle_params *p, double *F, double *sumF,
                   int32_t variant, double rel_tol, int64_t max_outer,
                   double *llh_out, double *llh_trace, int64_t trace_cap) {
    const size_t bytes = sizeof(double) * (size_t)n * (size_t)p->k;
    double *Fb = (double *)malloc(bytes);
    int64_t calls = 0;
    double LLHold;
    if (variant == 4) {
        LLHold = oracle_step(n, rowptr, col, p, F, sumF, NULL, Fb, NULL, NULL, NULL, 1, NULL, NULL);
        memcpy(F, Fb, bytes);
        if (llh_trace && calls < trace_cap) llh_trace[calls] = LLHold;
        ++calls;
    } else if (variant == 3) {
        LLHold = 0.0;
    } else {
        LLHold = oracle_llh(n, rowptr, col, p, F, sumF, NULL);
    }
    double last = LLHold;
    while (max_outer == 0 || calls < max_outer) {
        double newLLH = oracle_step(n, rowptr, col, p, F, sumF, NULL, Fb, NULL, NULL, NULL, 1, NULL, NULL);
        memcpy(F, Fb, bytes);
        if (llh_trace && calls < trace_cap) llh_trace[calls] = newLLH;
        ++calls;
        last = newLLH;
        if (fabs(1.0 - newLLH / LLHold) < rel_tol) break;
        LLHold = newLLH;
    }
    free(Fb);
    if (llh_out) *llh_out = (variant == 4) ? LLHold : last;
    return calls;
}

int32_t oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
