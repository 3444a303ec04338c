/*
 * mnem_oracle.c -- CPU restatement of cbg-ethz/mnem's EM hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call it.  Nothing under mnem_b200/ links or imports it; the
 * product path is the CUDA library and fails loudly when that library is missing.
 *
 * It restates, in plain C, what the reference computes in interpreted R plus 84 lines of C++.  The
 * reference cannot be executed in this environment (no R, no Rcpp/RcppEigen headers), so every function
 * below cites the reference lines it follows (paths relative to /root/reference).
 *
 * PARITY PINNING.  Pinned by the reference itself:
 *   - the five routines of src/mm.cpp (transClose_W/_Ins/_Del, maxCol_row, eigenMapMatMult's marshalling) against the
 *     reference's own file compiled UNMODIFIED into oracle/_ref/libmm_ref.so (oracle/Makefile, stub headers in
 *     oracle/rstub/; tests/test_ref_mm.py);
 *   - getLL / getAffinity against the 9 fitted objects of data/app.rda that carry `probs`
 *     (tests/golden/app_rda_golden.npz); transitive.reduction against the 45 shipped component graphs (they are its
 *     fixed points, R/mnems.r:374); getIC replayed on the shipped fits; transitive.closure against the roxygen example
 *     R/mnems.r:546-548.
 * The data contraction, doMean, scoreAdj's sum and the greedy search have NO reference-owned expected values (the
 * reference's only unit test is an RNG-dependent property and R cannot be run here): for those rows parity is
 * UNPINNED and the oracle is a reading of the source, cross-checked by an independent NumPy reading
 * (tests/test_oracle_selfcheck.py).
 *
 * Besides the plain restatement the file holds two harness aids (see "Decision replay"): a recorder that writes the
 * decisions of an oracle fit as a tape, and a replay that follows the tape of a device fit wherever the oracle's own
 * decision differs by less than the rounding error -- the way fits are compared at sizes where the reference's
 * decisions are at rounding level.
 *
 * Numerics.  R's sum/colSums/rowSums accumulate in `long double` (80-bit x87 on x86-64 builds) and round
 * once; `%*%` (reference BLAS) and Eigen's GEBP (SSE2, no FMA, one accumulator per output, depth <= 100)
 * accumulate in double in ascending index order.  acc_t below is that long-double accumulator; build
 * with -DORC_ACC_DOUBLE to get the behaviour of an R built with --disable-long-double (or arm64).
 *
 * Layout: every matrix is column-major like R.  Graphs are uint8 0/1 S x S, adj[i + j*S] = 1 <=> edge
 * i -> j (parent row, child column).  theta is int32[E] in 0..S, 0 = unattached.
 *
 * pert is int32[C]: a value in 1..S is the single S-gene knocked down in the cell; a value <= 0 is minus the
 * bit mask of the knocked-down S-genes (bit s-1 = S-gene s; 0 = none), i.e. a column of Rho = getRho(D)
 * (R/mnems_low.r:308-324) for a multi-knock-down cell ("1_3").  The Rho path then follows the reference:
 * effect rows are min(1, t(Rho) %*% adj1) (R/mnems_low.r:613-616, R/mnems.r:433-435) and doMean divides the
 * contribution of a cell with n > 1 knock-downs by n (R/mnems_low.r:896-904).  Masks need S <= 31.
 */
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef ORC_ACC_DOUBLE
typedef double acc_t;
#else
typedef long double acc_t;
#endif

#define ORC_MAXS 64

/* Inner parallelism for single fits at BASELINE scale (tests/golden/make_scale_golden.py): OpenMP threads split
 * INDEPENDENT outputs only -- cells of the E-step, E-gene rows of the two contractions, candidates of a greedy
 * iteration, rows of the mixture sums -- so every accumulator still receives the same addends in the same order
 * and results are bit-identical for any thread count (tests/test_oracle_properties.py).  Default 1 (serial). */
static int g_inner = 1;
void orc_set_inner_threads(int n) { g_inner = n < 1 ? 1 : n; }
int orc_get_inner_threads(void) { return g_inner; }
#ifdef _OPENMP
#define ORC_PAR_FOR _Pragma("omp parallel for schedule(static) num_threads(g_inner) if (g_inner > 1)")
#else
#define ORC_PAR_FOR
#endif

/* set of knocked-down S-genes of a cell as a bit mask (see the note on `pert` above) */
static inline uint32_t pert_mask(int32_t p) { return p > 0 ? 1u << (p - 1) : (uint32_t)(-(int64_t)p); }
/* adj1[c, j] of R/mnems_low.r:613-616: min(1, sum_s Rho[s,c] * A[s,j]) */
static inline int cell_effect(const uint8_t *A, int S, uint32_t pm, int j)
{
    for (int s = 0; s < S; s++)
        if (((pm >> s) & 1u) && A[s + j * S]) return 1;
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* src/mm.cpp                                                                                       */
/* ------------------------------------------------------------------------------------------------ */

/* src/mm.cpp:15-30 transClose_W.  px[i*nc+j] is (row j, col i): X[j,i] |= X[k,i] && X[j,k], loops k,i,j. */
void orc_trans_close_w(uint8_t *x, int n)
{
    for (int k = 0; k < n; k++)
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++)
                x[i * n + j] = (uint8_t)(x[i * n + j] || (x[i * n + k] && x[k * n + j]));
}

/* src/mm.cpp:50-65 transClose_Ins, u,v 1-based: X[i,j] |= X[i,u] && X[v,j] for every (i,j), in place. */
void orc_trans_close_ins(uint8_t *x, int n, int u, int v)
{
    int k = u - 1, l = v - 1;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++)
            x[j * n + i] = (uint8_t)(x[j * n + i] || (x[j * n + l] && x[k * n + i]));
}

/* src/mm.cpp:32-48 transClose_Del, u,v 1-based: X[u,v] = X[u,v] || exists k: X[u,k] && X[k,v]. */
void orc_trans_close_del(uint8_t *x, int n, int u, int v)
{
    int i = v - 1, j = u - 1;
    for (int k = 0; k < n; k++) {
        x[i * n + j] = (uint8_t)(x[i * n + j] || (x[i * n + k] && x[k * n + j]));
        if (x[i * n + j] == 1) break;
    }
}

/* src/mm.cpp:66-84 maxCol_row: per row the first strictly-greater maximum, 1-based.  Rows whose entries
 * are all NaN / -Inf leave the reference output uninitialised; the oracle returns 0 for them. */
void orc_max_col_row(const double *x, int nr, int nc, int32_t *out)
{
    for (int i = 0; i < nr; i++) {
        double best = -INFINITY;
        int32_t arg = 0;
        for (int j = 0; j < nc; j++)
            if (x[i + (size_t)j * nr] > best) { best = x[i + (size_t)j * nr]; arg = j + 1; }
        out[i] = arg;
    }
}

/* src/mm.cpp:8-13 eigenMapMatMult: C = A(m x p) * B(p x n), double, ascending-p sum from 0, no FMA. */
void orc_matmul(const double *A, const double *B, int m, int p, int n, double *Cout)
{
    for (int j = 0; j < n; j++)
        for (int i = 0; i < m; i++) {
            double s = 0.0;
            for (int l = 0; l < p; l++) {
                s += A[i + (size_t)l * m] * B[l + (size_t)j * p];
            }
            Cout[i + (size_t)j * m] = s;
        }
}

/* ------------------------------------------------------------------------------------------------ */
/* graph helpers: R/mnems_low.r:1199-1215 (mytc), R/mnems.r:511-533 (transitive.reduction)          */
/* ------------------------------------------------------------------------------------------------ */

static void set_diag(uint8_t *x, int n, uint8_t v) { for (int i = 0; i < n; i++) x[i * n + i] = v; }

/* mytc(x): diag(x) <- 1; transClose_W; binarise.  R/mnems_low.r:1199-1202,1211 */
void orc_mytc(uint8_t *x, int n) { set_diag(x, n, 1); orc_trans_close_w(x, n); }

/* mytc(x,u,v): diag 1, then Ins if x[u,v]==1 else Del.  R/mnems_low.r:1203-1210 */
void orc_mytc_uv(uint8_t *x, int n, int u, int v)
{
    set_diag(x, n, 1);
    if (x[(u - 1) + (v - 1) * n] == 1) orc_trans_close_ins(x, n, u, v);
    else orc_trans_close_del(x, n, u, v);
}

/* transitive.reduction: closure (diag 1), zero diag, then the in-place order-dependent triple loop.
 * R/mnems.r:517-532.  `type` is all zero for 0/1 input so the sign test is always true. */
void orc_trans_reduce(const uint8_t *g_in, int n, uint8_t *g)
{
    memcpy(g, g_in, (size_t)n * n);
    for (int i = 0; i < n * n; i++) g[i] = g[i] != 0;
    orc_mytc(g, n);
    set_diag(g, n, 0);
    for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++)
            if (g[x + y * n])
                for (int j = 0; j < n; j++)
                    if (g[y + j * n]) g[x + j * n] = 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* responsibilities and likelihood: R/mnems.r:1390-1441 (getAffinity), R/mnems_low.r:540-552 (getLL) */
/* ------------------------------------------------------------------------------------------------ */

static double rpow(double b, double y)
{   /* R_pow (arithmetic.c): x==1 or y==0 -> 1; y==2 -> x*x; else libm pow */
    if (b == 1.0 || y == 0.0) return 1.0;
    if (y == 2.0) return b * b;
    return pow(b, y);
}

/* getAffinity(x, affinity, norm=TRUE, logtype, mw, complete).  L is k x C, out is k x C.
 * affinity==0: soft path R/mnems.r:1419-1437.  affinity==1: hard path :1393-1417 (complete must be 0 --
 * the reference's complete branch there reads `y` before it is defined and errors). */
void orc_get_affinity(const double *L, int k, int C, const double *mw_in, int affinity, int complete,
                      double logtype, double *out)
{
    double mwbuf[256];
    const double *mw = mw_in;
    if (!mw) { for (int i = 0; i < k; i++) mwbuf[i] = 1.0 / k; mw = mwbuf; }   /* :1392 */
    if (affinity == 1) {
        ORC_PAR_FOR
        for (int c = 0; c < C; c++) {
            double mx = -INFINITY;
            for (int i = 0; i < k; i++) {
                out[i + (size_t)c * k] = rpow(logtype, L[i + (size_t)c * k]) * mw[i];       /* :1403-1404 */
                if (out[i + (size_t)c * k] > mx) mx = out[i + (size_t)c * k];
            }
            for (int i = 0; i < k; i++) {
                double y = out[i + (size_t)c * k];
                if (y < mx) y = 0.0;                                                    /* :1405-1411 */
                if (y != 0.0) y = 1.0;                                                  /* :1415-1417 */
                if (isnan(y)) y = 0.0;                                                  /* :1439 */
                out[i + (size_t)c * k] = y;
            }
        }
        return;
    }
    if (k == 1) { for (int c = 0; c < C; c++) out[c] = 1.0; return; }                    /* :1435-1437 */
    int shift = 0;
    if (complete) {                                                                      /* :1422 any(y > 0) */
        for (size_t i = 0; i < (size_t)k * C; i++) if (L[i] > 0) { shift = 1; break; }
    }
    double shrink = log(ldexp(1.0, 1023)) / log(logtype);                                /* :1425-1426 */
    ORC_PAR_FOR
    for (int c = 0; c < C; c++) {
        const double *x = L + (size_t)c * k;
        double *y = out + (size_t)c * k;
        double xmax = -INFINITY;
        if (shift) for (int i = 0; i < k; i++) if (x[i] > xmax) xmax = x[i];
        acc_t cs = 0;
        for (int i = 0; i < k; i++) {
            double v = x[i];
            if (shift) v = v - (xmax - shrink / k);                                      /* :1427 */
            v = rpow(logtype, v);                                                        /* :1431 */
            v = v * mw[i];                                                               /* :1432 */
            y[i] = v;
            cs += v;                                                                     /* colSums :1433 */
        }
        double csd = (double)cs;
        for (int i = 0; i < k; i++) {
            double v = y[i] / csd;
            if (isnan(v)) v = 0.0;                                                       /* :1439 */
            y[i] = v;
        }
    }
}

/* getLL.  R/mnems_low.r:540-552. */
double orc_get_ll(const double *L, int k, int C, const double *mw_in, int complete, double logtype)
{
    double mwbuf[256];
    const double *mw = mw_in;
    if (!mw) { for (int i = 0; i < k; i++) mwbuf[i] = 1.0 / k; mw = mwbuf; }
    acc_t l = 0;
    if (complete) {                                                                       /* :542-545 */
        double *Z = (double *)malloc(sizeof(double) * (size_t)k * C);
        orc_get_affinity(L, k, C, mw, 0, 1, logtype, Z);
        double lmw[256];
        for (int i = 0; i < k; i++) lmw[i] = log(mw[i]) / log(logtype);
        double *term = (double *)malloc(sizeof(double) * (size_t)C);
        ORC_PAR_FOR
        for (int c = 0; c < C; c++) {
            acc_t s = 0;
            for (int i = 0; i < k; i++) s += Z[i + (size_t)c * k] * (L[i + (size_t)c * k] + lmw[i]);
            term[c] = (double)s;
        }
        for (int c = 0; c < C; c++) l += term[c];            /* sum(): one sequential long-double accumulator */
        free(term);
        free(Z);
    } else {                                                                              /* :547-549 */
        double lb = log(logtype);
        double *term = (double *)malloc(sizeof(double) * (size_t)C);
        ORC_PAR_FOR
        for (int c = 0; c < C; c++) {
            double s = 0.0;                                 /* rep(1,k) %*% x : double, ascending k */
            for (int i = 0; i < k; i++) s += 1.0 * (rpow(logtype, L[i + (size_t)c * k]) * mw[i]);
            term[c] = log(s) / lb;
        }
        for (int c = 0; c < C; c++) l += term[c];
        free(term);
    }
    return (double)l;
}

/* mw <- apply(getAffinity(L, mw), 1, sum); mw <- mw/sum(mw); NA -> 1/k.  R/mnems_low.r:591-595, 653-657,
 * R/mnems.r:1936-1940, 2113-2121 (the NA guard exists only at :595 and :1940, selected by na_guard). */
static void mw_update_a(const double *L, int k, int C, double *mw, int affinity, int complete, double logtype, int na_guard,
                      double *scratch)
{
    orc_get_affinity(L, k, C, mw, affinity, complete, logtype, scratch);
    acc_t tot = 0;
    double rs[256];
    ORC_PAR_FOR
    for (int i = 0; i < k; i++) {
        acc_t s = 0;
        for (int c = 0; c < C; c++) s += scratch[i + (size_t)c * k];
        rs[i] = (double)s;
    }
    for (int i = 0; i < k; i++) tot += rs[i];
    int bad = 0;
    for (int i = 0; i < k; i++) { mw[i] = rs[i] / (double)tot; if (isnan(mw[i])) bad = 1; }
    if (bad && na_guard) for (int i = 0; i < k; i++) mw[i] = 1.0 / k;
}

static void mw_update(const double *L, int k, int C, double *mw, int complete, double logtype, int na_guard,
                      double *scratch)
{
    mw_update_a(L, k, C, mw, 0, complete, logtype, na_guard, scratch);
}

void orc_mw_update(const double *L, int k, int C, double *mw, int complete, double logtype, int na_guard)
{
    double *scratch = (double *)malloc(sizeof(double) * (size_t)k * C);
    mw_update(L, k, C, mw, complete, logtype, na_guard, scratch);
    free(scratch);
}

/* ------------------------------------------------------------------------------------------------ */
/* M-step sufficient statistics: doMean, Rho path.  R/mnems_low.r:893-906 reached from nem() R/mnems.r:75-77, */
/* 101-107 (nem fills Rho <- getRho(D) when NULL, so the rowsum/global-divisor branch :908-911 is never  */
/* taken on the EM path).  R[e,s] = sum_c (D[e,c]*w[c]) * Rho[s,c], ascending c, double (BLAS dgemm).  */
/* The caller block R/mnems.r:1986-2006 drops zero-weight cells and appends S all-zero cells: the sums  */
/* are unchanged.  weights == NULL: plain D (k == 1 path, R/mnems.r:1851).                               */
/* ------------------------------------------------------------------------------------------------ */
/* rows [e0, e1) of nw collapsed matrices at once: weight q of cell c is w[q + c*stride] (w == NULL: plain D), output q
 * is R + q*E*S.  One pass over D; every R[e,s] still receives its addends in ascending cell order. */
static void do_mean_slice(const double *D, int E, int C, const int32_t *pert, int S, int nw, const double *w, size_t stride,
                          double *R, int e0, int e1)
{
    for (int q = 0; q < nw; q++)
        for (int s = 0; s < S; s++) memset(R + ((size_t)q * S + s) * E + e0, 0, sizeof(double) * (size_t)(e1 - e0));
    for (int c = 0; c < C; c++) {
        const double *d = D + (size_t)c * E;
        const uint32_t pm = pert_mask(pert[c]);
        const int n = __builtin_popcount(pm);
        for (int q = 0; q < nw; q++) {
            const double wc = w ? w[q + (size_t)c * stride] : 1.0;
            if (w && wc == 0.0) continue;
            double *Rq = R + (size_t)q * S * E;
            if (n == 1) {
                double *r = Rq + (size_t)__builtin_ctz(pm) * E;
                if (w) {
                    for (int e = e0; e < e1; e++) r[e] += d[e] * wc;   /* built with -ffp-contract=off: mul then add */
                } else {
                    for (int e = e0; e < e1; e++) r[e] += d[e];
                }
            } else if (n > 1) {
                /* Rho column normalised to sum 1 (:897-902): each knocked-down S-gene receives (D*w) * (1/n) */
                const double rho = 1.0 / n;
                for (int s = 0; s < S; s++)
                    if ((pm >> s) & 1u) {
                        double *r = Rq + (size_t)s * E;
                        for (int e = e0; e < e1; e++) r[e] += (w ? d[e] * wc : d[e]) * rho;
                    }
            }
        }
    }
}

/* nw weight vectors (rows of a k x C column-major matrix when stride = k), one pass; threads own E-gene row ranges */
static void do_mean_multi(const double *D, int E, int C, const int32_t *pert, int S, int nw, const double *w, size_t stride,
                          double *R)
{
    int nt = g_inner < 1 ? 1 : g_inner;
    if (nt > (E + 7) / 8) nt = (E + 7) / 8;
    if (nt <= 1) { do_mean_slice(D, E, C, pert, S, nw, w, stride, R, 0, E); return; }
    int t;
#ifdef _OPENMP
#pragma omp parallel for schedule(static, 1) num_threads(nt)
#endif
    for (t = 0; t < nt; t++) {
        const int per = ((E + nt - 1) / nt + 7) / 8 * 8;
        const int e0 = t * per, e1 = e0 + per < E ? e0 + per : E;
        if (e0 < e1) do_mean_slice(D, E, C, pert, S, nw, w, stride, R, e0, e1);
    }
}

void orc_do_mean(const double *D, int E, int C, const int32_t *pert, int S, const double *w, double *R)
{
    do_mean_multi(D, E, C, pert, S, 1, w, 1, R);
}

/* ------------------------------------------------------------------------------------------------ */
/* scoreAdj on the collapsed E x S matrix.  R/mnems.r:416-458 with method="llr", marginal=FALSE,       */
/* Rho = identity (R/mnems.r:104), weights = NULL: W = R %*% adj (Eigen), theta = first argmax over     */
/* cbind(0, W) minus 1 (:447-450), score = sum(rowMaxs(cbind(0, W))) (:452-454, long-double sum).       */
/* ------------------------------------------------------------------------------------------------ */
static void score_cols(const double *R, int E, int S, const uint8_t *adj, const uint8_t *colsel, double *W)
{
    for (int j = 0; j < S; j++) {
        if (colsel && !colsel[j]) continue;
        double *w = W + (size_t)j * E;
        for (int e = 0; e < E; e++) w[e] = 0.0;
        for (int s = 0; s < S; s++)
            if (adj[s + j * S]) {
                const double *r = R + (size_t)s * E;
                for (int e = 0; e < E; e++) w[e] += r[e];        /* product with 1.0 is exact */
            }
    }
}

static double score_total(const double *W, int E, int S, int32_t *theta)
{
    acc_t tot = 0;
    for (int e = 0; e < E; e++) {
        double best = 0.0;
        int32_t arg = 0;
        for (int j = 0; j < S; j++) {
            double v = W[e + (size_t)j * E];
            if (v > best) { best = v; arg = j + 1; }
        }
        if (theta) theta[e] = arg;
        tot += best;
    }
    return (double)tot;
}

/* marginal = TRUE, R/mnems.r:455-457: pscore <- logtype^cbind(0, score); sum(log(rowSums(pscore))/log(logtype)).
 * rowSums and sum accumulate in long double; theta (dotopo) is the same argmax as in the max variant (:447-450). */
static double score_total_marginal(const double *W, int E, int S, double logtype, int32_t *theta)
{
    acc_t tot = 0;
    const double lb = log(logtype);
    for (int e = 0; e < E; e++) {
        double best = 0.0;
        int32_t arg = 0;
        acc_t rs = rpow(logtype, 0.0);
        for (int j = 0; j < S; j++) {
            double v = W[e + (size_t)j * E];
            if (v > best) { best = v; arg = j + 1; }
            rs += rpow(logtype, v);
        }
        if (theta) theta[e] = arg;
        tot += log((double)rs) / lb;
    }
    return (double)tot;
}

/* Public: scoreAdj(D=R, adj, trans.close, marginal, logtype, dotopo=TRUE).  W_out (E x S) and theta may be NULL. */
double orc_score_adj_m(const double *R, int E, int S, const uint8_t *adj_in, int trans_close, int marginal,
                       double logtype, int32_t *theta, double *W_out)
{
    uint8_t adj[ORC_MAXS * ORC_MAXS];
    memcpy(adj, adj_in, (size_t)S * S);
    if (trans_close) orc_mytc(adj, S);                                                   /* :428-430 */
    double *W = W_out ? W_out : (double *)malloc(sizeof(double) * (size_t)E * S);
    score_cols(R, E, S, adj, NULL, W);
    double sc = marginal ? score_total_marginal(W, E, S, logtype, theta) : score_total(W, E, S, theta);
    if (!W_out) free(W);
    return sc;
}
double orc_score_adj(const double *R, int E, int S, const uint8_t *adj_in, int trans_close, int32_t *theta,
                     double *W_out)
{
    return orc_score_adj_m(R, E, S, adj_in, trans_close, 0, 2.0, theta, W_out);
}

/* ------------------------------------------------------------------------------------------------ */
/* neighbour generators R/mnems_low.r:1025-1084, combined as unique(c(ins, rev, del)) R/mnems.r:222-224 */
/* ------------------------------------------------------------------------------------------------ */
/* Writes at most 3*S*S matrices (S*S bytes each) into `models`; returns the number written BEFORE the
 * keep-first dedupe in *n_raw and the number after it as the return value.  kind[i]: 0 ins, 1 rev, 2 del. */
int orc_neighbours(const uint8_t *P, int S, uint8_t *models, int8_t *kind, int *n_raw)
{
    const int SS = S * S;
    int n = 0;
    uint8_t T[ORC_MAXS * ORC_MAXS];
    /* get.ins.fast :1048-1068 */
    for (int idx = 0; idx < SS; idx++)
        if (P[idx] == 0) {
            uint8_t *m = models + (size_t)n * SS;
            memcpy(m, P, SS);
            m[idx] = 1;
            orc_mytc_uv(m, S, idx % S + 1, idx / S + 1);
            kind[n++] = 0;
        }
    /* get.rev.tc :1025-1046 */
    orc_trans_reduce(P, S, T);
    for (int c = 0; c < S; c++)
        for (int r = 0; r < S; r++)
            if (T[r + c * S] + T[c + r * S] == 1) {
                uint8_t *m = models + (size_t)n * SS;
                memcpy(m, P, SS);
                m[r + c * S] = (uint8_t)(1 - m[r + c * S]);
                if (r != c) m[c + r * S] = (uint8_t)(1 - m[c + r * S]);
                set_diag(m, S, 1);
                if (m[r + c * S] == 1) orc_mytc_uv(m, S, r + 1, c + 1);
                else orc_mytc_uv(m, S, c + 1, r + 1);
                kind[n++] = 1;
            }
    /* get.del.tc :1070-1084 */
    {
        uint8_t P0[ORC_MAXS * ORC_MAXS];
        memcpy(P0, P, SS);
        for (int i = 0; i < S; i++) P0[i + i * S] = (uint8_t)(P0[i + i * S] ? P0[i + i * S] - 1 : 0);
        orc_trans_reduce(P0, S, T);
        for (int idx = 0; idx < SS; idx++)
            if (T[idx] == 1) {
                uint8_t *m = models + (size_t)n * SS;
                memcpy(m, P0, SS);
                m[idx] = 0;
                set_diag(m, S, 1);
                kind[n++] = 2;
            }
    }
    if (n_raw) *n_raw = n;
    /* unique(): keep first occurrence, R/mnems.r:222 */
    int nu = 0;
    for (int i = 0; i < n; i++) {
        int dup = 0;
        for (int j = 0; j < nu && !dup; j++)
            if (memcmp(models + (size_t)i * SS, models + (size_t)j * SS, SS) == 0) dup = 1;
        if (!dup) {
            if (nu != i) { memmove(models + (size_t)nu * SS, models + (size_t)i * SS, SS); kind[nu] = kind[i]; }
            nu++;
        }
    }
    return nu;
}

/* ------------------------------------------------------------------------------------------------ */
/* Decision replay ("advice").  TEST HARNESS, not part of the reference's algorithm.                    */
/*                                                                                                      */
/* Every graph the EM returns is the outcome of comparisons between floating-point sums, and at scale a  */
/* few per cent of them are closer than the rounding error of ANY summation order (the greedy search     */
/* keeps accepting moves that improve a score of 4e6 by 1e-9, see n_dec_tight).  Two correct             */
/* implementations then legitimately part ways.  To still compare a device fit with this oracle decision */
/* by decision, the device records every such decision on a tape (include/mnemcu.h, mnemcu_em_opts::tape)*/
/* and the oracle replays the fit in ITS OWN arithmetic: wherever its decision differs from the tape's   */
/* it measures the margin in its own numbers; a margin <= tol means "decided by rounding" and the tape's  */
/* choice is taken (counted in n_forced), a larger one is a genuine disagreement (n_hard; the replay then */
/* stops listening to the tape).  With n_hard == 0 both fits walk through the same graphs, and the final  */
/* results are compared exactly (graphs, theta) and to 1e-9 (likelihoods) by the caller.                  */
/* ------------------------------------------------------------------------------------------------ */
enum { ADV_MOVE = 0, ADV_STOP, ADV_PICK, ADV_THETA, ADV_INIT_THETA, ADV_INIT_BEST, ADV_ACCEPT, ADV_BEST, ADV_KINDS };
typedef struct {
    double tol;                  /* relative margin below which a score / weight comparison counts as rounding noise */
    double tol_ll;               /* same for log-likelihood comparisons */
    int T;                       /* inner iterations of the theta-free getProbs on tape */
    const int32_t *init_theta;   /* [T][k*E] in 0..S */
    const int32_t *init_best;    /* [T] */
    int n_em;                    /* EM iterations on tape */
    const int64_t *move_off;     /* [n_em*k*2 + 1] first move of search (iter, comp, restart) in move_rows */
    const uint32_t *move_rows;   /* [moves][S] row masks (bit j of row s = edge s -> j, diagonal set) after each accepted move */
    const int32_t *pick;         /* [n_em*k] */
    const int32_t *theta;        /* [n_em][k*E] */
    const int32_t *accept;       /* [n_em] */
    const int32_t *best_at;      /* [n_em + 1]: best result replaced after `count` iterations */
    /* report */
    long n_checked[ADV_KINDS], n_forced[ADV_KINDS], n_hard[ADV_KINDS];
    double max_forced[ADV_KINDS];     /* largest margin that was overridden */
    double min_hard[ADV_KINDS];       /* smallest margin that was NOT (first evidence of a real disagreement) */
    int diverged;
    char first_hard[256];
} orc_advice;

static __thread orc_advice *g_adv = NULL;           /* active replay of the calling thread */
static __thread const uint32_t *g_adv_rows = NULL;  /* moves of the search about to run */
static __thread int g_adv_nmoves = 0;
static __thread double *g_adv_W = NULL;             /* if set, nem_greedy leaves the subweights of its final graph here */
static __thread int g_adv_probs = 0;                /* the next getProbs call is the theta-free one on tape */
static __thread int g_nullcomp = 0;                 /* getProbs(nullcomp = TRUE): the last component's log ratios are 0 */

/* The oracle can also WRITE a tape of its own fit, in the device's record format (include/mnemcu.h).  Used to test the
 * replay on the CPU: the tape of the double-accumulator build replayed by the long-double build is the same situation
 * as a device tape (two correct implementations that round differently). */
static __thread int32_t *g_rec = NULL;
static __thread int64_t g_rec_n = 0, g_rec_cap = 0;
static __thread int g_rec_start = 0, g_rec_iter = 0, g_rec_comp = 0, g_rec_restart = 0, g_rec_probs = 0;
void orc_set_recorder(int32_t *buf, int64_t cap, int start) { g_rec = buf; g_rec_cap = cap; g_rec_n = 0; g_rec_start = start; }
int64_t orc_recorder_words(void) { return g_rec_n; }
static void rec_w(int32_t w) { if (g_rec) { if (g_rec_n < g_rec_cap) g_rec[g_rec_n] = w; g_rec_n++; } }
static void rec_theta(int code, int idx, const int32_t *th, size_t n, int S, int coded_null_last)
{
    if (!g_rec) return;
    rec_w(code); rec_w(g_rec_start); rec_w(idx);
    for (size_t i = 0; i < n; i++) rec_w(coded_null_last && th[i] == S + 1 ? 0 : th[i]);
}

static int adv_on(void) { return g_adv && !g_adv->diverged; }
static void adv_forced(int kind, double mg)
{
    g_adv->n_forced[kind]++;
    if (mg > g_adv->max_forced[kind]) g_adv->max_forced[kind] = mg;
}
static void adv_hard(int kind, double mg, const char *what, int a, int b, int c)
{
    g_adv->n_hard[kind]++;
    if (mg < g_adv->min_hard[kind]) g_adv->min_hard[kind] = mg;
    if (!g_adv->diverged)
        snprintf(g_adv->first_hard, sizeof g_adv->first_hard, "%s (at %d, %d, %d): margin %.3e > tolerance", what, a, b, c, mg);
    g_adv->diverged = 1;
}
/* theta of one component against the tape: W is E x ncol column-major (subweights), own[e] / tape[e] in 0..S with
 * 0 = the null column (value 0).  Entries that differ are overridden when the two weights are closer than
 * tol * sum_j |W[e, j]|.  Returns the number of overridden entries. */
static int adv_theta(int kind, const double *W, int E, int S, int32_t *own, const int32_t *tape, int where_a, int where_b)
{
    int nf = 0;
    for (int e = 0; e < E && adv_on(); e++) {
        g_adv->n_checked[kind]++;
        if (own[e] == tape[e]) continue;
        double sc = 0;
        for (int j = 0; j < S; j++) sc += fabs(W[e + (size_t)j * E]);
        const double wa = own[e] == 0 ? 0.0 : W[e + (size_t)(own[e] - 1) * E];
        const double wb = (tape[e] < 1 || tape[e] > S) ? 0.0 : W[e + (size_t)(tape[e] - 1) * E];
        const double mg = fabs(wa - wb) / (sc > 1e-300 ? sc : 1e-300);
        if (tape[e] >= 0 && tape[e] <= S && mg <= g_adv->tol) { adv_forced(kind, mg); own[e] = tape[e]; nf++; }
        else adv_hard(kind, mg, kind == ADV_THETA ? "theta after the M-step" : "theta of the initial E-step", where_a, where_b, e);
    }
    return nf;
}

/* ------------------------------------------------------------------------------------------------ */
/* nem(search="greedy", runs=1, domean already applied): R/mnems.r:111-142, 203-307, 362-379           */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    double score;        /* nem$score = oldscore */
    double max_trace;    /* max(nem$scores) */
    int n_iter;          /* accepted moves */
    long n_scored;       /* scoreAdj calls inside the loop (unique candidates) */
    double min_margin;   /* smallest |best - runner-up or old| seen in a decision, relative; diagnostic */
    long n_raw;          /* candidates generated before unique() */
    long n_dec;          /* greedy decisions taken (one argmax + accept/stop test per iteration) */
    long n_dec_tight;    /* ... of which the winner beat the runner-up or the old score by less than 1e-13 relative (but not
                          * by exactly 0): decided by the last bits of the sums, i.e. by the summation order / accumulator width */
} orc_nem_info;

void orc_nem_greedy_m(const double *R, int E, int S, const uint8_t *start, int marginal, double logtype,
                      uint8_t *adj_out, int32_t *theta_out, orc_nem_info *info);
void orc_nem_greedy(const double *R, int E, int S, const uint8_t *start, uint8_t *adj_out, int32_t *theta_out,
                    orc_nem_info *info)
{
    orc_nem_greedy_m(R, E, S, start, 0, 2.0, adj_out, theta_out, info);
}

/* marginal != 0: every scoreAdj call of the search carries marginal = TRUE (R/mnems.r:135, 208, 363) */
void orc_nem_greedy_m(const double *R, int E, int S, const uint8_t *start, int marginal, double logtype,
                      uint8_t *adj_out, int32_t *theta_out, orc_nem_info *info)
{
    const int SS = S * S;
    uint8_t better[ORC_MAXS * ORC_MAXS], tmp[ORC_MAXS * ORC_MAXS];
    if (start) memcpy(better, start, SS); else memset(better, 0, SS);
    for (int i = 0; i < SS; i++) better[i] = better[i] != 0;
    set_diag(better, S, 1);                                                      /* :133, NOT closed */
    double *P = (double *)malloc(sizeof(double) * (size_t)E * S);                 /* subweights of `better` */
    double *Wc = (double *)malloc(sizeof(double) * (size_t)E * S);
    double *Pbest = (double *)malloc(sizeof(double) * (size_t)E * S);
    uint8_t *models = (uint8_t *)malloc((size_t)3 * SS * SS + SS);
    int8_t *kind = (int8_t *)malloc((size_t)3 * SS + 1);
    int haveP = 0;
    int64_t rec_hdr = -1;
    if (g_rec) { rec_w(2); rec_w(g_rec_start); rec_w(g_rec_iter); rec_w(g_rec_comp); rec_w(g_rec_restart); rec_hdr = g_rec_n; rec_w(0); }
    double oldscore = orc_score_adj_m(R, E, S, better, 1, marginal, logtype, NULL, NULL);   /* :135-141 closes inside */
    double maxtrace = oldscore;
    long nscored = 0, nraw = 0, ndec = 0, ntight = 0;
    int niter = 0;
    double minmargin = INFINITY;
    for (;;) {
        int nr1 = 0;
        int nm = orc_neighbours(better, S, models, kind, &nr1);
        nraw += nr1;
        if (nm == 0) break;
        double bests = -INFINITY, second = -INFINITY;
        int bestidx = -1;
        /* scoreAdj of one candidate into Wbuf; returns the score (NA -> 0, :235) */
#define ORC_SCORE_CAND(m_, Wbuf_, sc_)                                                                       \
        do {                                                                                                 \
            const uint8_t *cand_ = models + (size_t)(m_) * SS;                                               \
            if (!haveP) {                                                        /* :439-440 */              \
                score_cols(R, E, S, cand_, NULL, (Wbuf_));                                                   \
            } else {                                                             /* :441-446 */              \
                uint8_t chg_[ORC_MAXS];                                                                      \
                for (int j = 0; j < S; j++) {                                                                \
                    chg_[j] = 0;                                                                             \
                    for (int s_ = 0; s_ < S; s_++) if (cand_[s_ + j * S] != better[s_ + j * S]) { chg_[j] = 1; break; } \
                }                                                                                            \
                memcpy((Wbuf_), P, sizeof(double) * (size_t)E * S);                                          \
                score_cols(R, E, S, cand_, chg_, (Wbuf_));                                                   \
            }                                                                                                \
            (sc_) = marginal ? score_total_marginal((Wbuf_), E, S, logtype, NULL) : score_total((Wbuf_), E, S, NULL); \
            if (isnan(sc_)) (sc_) = 0.0;                                         /* :235 */                  \
        } while (0)
        double *adv_scs = NULL;
        if ((g_inner > 1 && nm > 1) || adv_on()) {
            /* candidates are independent: score them on g_inner threads, then take which.max in candidate order */
            double *scs = (double *)malloc(sizeof(double) * (size_t)nm);
#ifdef _OPENMP
#pragma omp parallel num_threads(g_inner)
#endif
            {
                double *Wt = (double *)malloc(sizeof(double) * (size_t)E * S);
                int m;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 8)
#endif
                for (m = 0; m < nm; m++) { double sc; ORC_SCORE_CAND(m, Wt, sc); scs[m] = sc; }
                free(Wt);
            }
            for (int m = 0; m < nm; m++) {
                const double sc = scs[m];
                nscored++;
                if (sc > bests) { second = bests; bests = sc; bestidx = m; }     /* which.max: first */
                else if (sc > second && sc < bests) second = sc;
            }
            if (adv_on()) adv_scs = scs; else free(scs);
            if (!adv_scs) { double sc; ORC_SCORE_CAND(bestidx, Pbest, sc); (void)sc; }   /* subweights of the winner */
        } else {
            for (int m = 0; m < nm; m++) {
                double sc;
                ORC_SCORE_CAND(m, Wc, sc);
                nscored++;
                if (sc > bests) {                                                /* which.max: first */
                    second = bests; bests = sc; bestidx = m;
                    memcpy(Pbest, Wc, sizeof(double) * (size_t)E * S);
                } else if (sc > second && sc < bests) second = sc;
            }
        }
        const uint8_t *best = models + (size_t)bestidx * SS;
        int nb = 0, nc = 0;
        for (int i = 0; i < SS; i++) { nb += better[i] == 1; nc += best[i] == 1; }
        double scale = fabs(bests) > 1e-300 ? fabs(bests) : 1e-300;   /* margins relative to the score itself */
        int force_accept = 0, force_stop = 0;
        if (adv_scs) {
            /* replay: the tape's decision for this iteration against the oracle's own (see orc_advice) */
            const int accept_own = bests > oldscore || (bests == oldscore && nb > nc);
            if (niter < g_adv_nmoves) {
                const uint32_t *rows = g_adv_rows + (size_t)niter * S;
                int mstar = -1;
                for (int m = 0; m < nm && mstar < 0; m++) {
                    const uint8_t *cand = models + (size_t)m * SS;
                    int same = 1;
                    for (int j = 0; j < S && same; j++)
                        for (int s_ = 0; s_ < S; s_++)
                            if ((cand[s_ + j * S] != 0) != (int)((rows[s_] >> j) & 1u)) { same = 0; break; }
                    if (same) mstar = m;
                }
                g_adv->n_checked[ADV_MOVE]++;
                if (mstar < 0) adv_hard(ADV_MOVE, INFINITY, "the taped move is no neighbour of the current graph", niter, nm, 0);
                else if (!(mstar == bestidx && accept_own)) {
                    const double m1 = (bests - adv_scs[mstar]) / scale, m2 = (oldscore - adv_scs[mstar]) / scale;
                    const double mg = m1 > m2 ? (m1 > 0 ? m1 : 0) : (m2 > 0 ? m2 : 0);
                    if (mg <= g_adv->tol) {
                        adv_forced(ADV_MOVE, mg);
                        if (getenv("ORC_TRACE_MARGIN"))
                            fprintf(stderr, "forced move: iter %d own idx %d (%.17g, accept %d) tape idx %d (%.17g) old %.17g\n", niter,
                                    bestidx, bests, accept_own, mstar, adv_scs[mstar], oldscore);
                        bestidx = mstar; bests = adv_scs[mstar]; best = models + (size_t)bestidx * SS; force_accept = 1;
                    } else adv_hard(ADV_MOVE, mg, "greedy move", niter, bestidx, mstar);
                }
            } else {
                g_adv->n_checked[ADV_STOP]++;
                if (accept_own) {
                    const double mg = (bests - oldscore) / scale;
                    if (mg <= g_adv->tol) { adv_forced(ADV_STOP, mg); force_stop = 1; }
                    else adv_hard(ADV_STOP, mg, "greedy search stops on the tape but not here", niter, bestidx, 0);
                }
            }
            { double sc; ORC_SCORE_CAND(bestidx, Pbest, sc); (void)sc; }          /* subweights of the move taken */
            free(adv_scs);
        }
#undef ORC_SCORE_CAND
        if (bests != oldscore && fabs(bests - oldscore) / scale < minmargin) minmargin = fabs(bests - oldscore) / scale;
        if (isfinite(second) && (bests - second) / scale < minmargin) minmargin = (bests - second) / scale;
        ndec++;
        if ((bests != oldscore && fabs(bests - oldscore) / scale < 1e-13) ||
            (isfinite(second) && (bests - second) / scale < 1e-13)) ntight++;
        if (getenv("ORC_TRACE_MARGIN") && ((bests != oldscore && fabs(bests - oldscore) / scale < 1e-13) ||
                                           (isfinite(second) && (bests - second) / scale < 1e-13)))
            fprintf(stderr, "near-tie: greedy iter %d nm %d best %.17g (idx %d kind %d) second %.17g old %.17g\n", niter, nm,
                    bests, bestidx, (int)kind[bestidx], second, oldscore);
        if (!force_stop && (force_accept || bests > oldscore || (bests == oldscore && nb > nc))) {   /* :279-286 */
            memcpy(tmp, best, SS); memcpy(better, tmp, SS);
            if (g_rec)
                for (int s_ = 0; s_ < S; s_++) {
                    uint32_t rw = 0;
                    for (int j = 0; j < S; j++) if (better[s_ + j * S]) rw |= 1u << j;
                    rec_w((int32_t)rw);
                }
            oldscore = bests;
            if (oldscore > maxtrace) maxtrace = oldscore;
            memcpy(P, Pbest, sizeof(double) * (size_t)E * S);
            haveP = 1;
            niter++;
        } else break;                                                            /* :287-289 */
    }
    /* :362-369 theta from the closed final graph, :374 reduction */
    if (g_rec && rec_hdr >= 0 && rec_hdr < g_rec_cap) g_rec[rec_hdr] = niter;
    orc_score_adj_m(R, E, S, better, 1, marginal, logtype, theta_out, g_adv_W);
    orc_trans_reduce(better, S, adj_out);
    if (info) { info->score = oldscore; info->max_trace = maxtrace; info->n_iter = niter; info->n_scored = nscored;
                info->min_margin = minmargin; info->n_raw = nraw; info->n_dec = ndec; info->n_dec_tight = ntight; }
    free(P); free(Wc); free(Pbest); free(models); free(kind);
}

/* replay of theta inside the theta-free getProbs: th[e] in 1..S+1 (S+1 = the appended null column), tape in 0..S */
static void adv_init_theta(const double *W, int E, int S, int32_t *th, const int32_t *tape, int count, int comp)
{
    int32_t *own = (int32_t *)malloc(sizeof(int32_t) * E);
    for (int e = 0; e < E; e++) own[e] = th[e] == S + 1 ? 0 : th[e];
    if (adv_theta(ADV_INIT_THETA, W, E, S, own, tape, count, comp) > 0)
        for (int e = 0; e < E; e++) th[e] = own[e] == 0 ? S + 1 : own[e];
    free(own);
}

/* ------------------------------------------------------------------------------------------------ */
/* E-step: getProbs.  R/mnems_low.r:576-663                                                          */
/* ------------------------------------------------------------------------------------------------ */
/* phi: k graphs (S*S each; any form -- closed inside, :610).  theta: k*E in 0..S or NULL (argmax path
 * :618-619).  L_in: k x C (the `probs` argument).  Outputs: L_out k x C (bestprobs), mw_out[k], returns ll
 * (llold of the last inner iteration, :661-662).  faithful_cost != 0 recomputes the contraction in every
 * inner iteration as the reference does (identical values when theta is given). */
double orc_get_probs_a(const double *D, int E, int C, const int32_t *pert, int S, int k, const uint8_t *phi,
                       const int32_t *theta, const double *L_in, const double *mw_in, int affinity, int complete,
                       double logtype, int faithful_cost, double *L_out, double *mw_out, int32_t *theta_used)
{
    const int SS = S * S;
    size_t kc = (size_t)k * C;
    double *probs = (double *)malloc(sizeof(double) * kc);
    double *probs0 = (double *)malloc(sizeof(double) * kc);
    double *best = (double *)malloc(sizeof(double) * kc);
    double *post = (double *)malloc(sizeof(double) * kc);
    double *scratch = (double *)malloc(sizeof(double) * kc);
    uint8_t *clo = (uint8_t *)malloc((size_t)k * SS);
    int32_t *th = (int32_t *)malloc(sizeof(int32_t) * (size_t)k * E);
    double *Wtmp = theta ? NULL : (double *)malloc(sizeof(double) * (size_t)E * (S + 1));
    memcpy(probs, L_in, sizeof(double) * kc);
    memcpy(best, L_in, sizeof(double) * kc);                                      /* :581 */
    double mw[256];
    for (int i = 0; i < k; i++) mw[i] = mw_in ? mw_in[i] : 1.0 / k;
    int max_count = C <= 100 ? 100 : 10;                                          /* :584-588 */
    double llold = -INFINITY;
    mw_update_a(probs, k, C, mw, affinity, complete, logtype, 1, scratch);       /* :591-595 */
    for (int i = 0; i < k; i++) { memcpy(clo + (size_t)i * SS, phi + (size_t)i * SS, SS);
        for (int q = 0; q < SS; q++) clo[(size_t)i * SS + q] = clo[(size_t)i * SS + q] != 0;
        orc_mytc(clo + (size_t)i * SS, S); }                                     /* :610 */
    int have0 = 0;
    const int adv_probs = g_adv_probs;                     /* replay: this is the theta-free call the tape describes */
    g_adv_probs = 0;
    const int rec_probs = g_rec_probs && !theta;
    g_rec_probs = 0;
    for (int count = 0; count < max_count; count++) {                            /* converged = -Inf: never stops early */
        const int adv_here = adv_probs && !theta && adv_on() && count < g_adv->T;
        if (!theta) orc_get_affinity(probs, k, C, mw, affinity, complete, logtype, post);   /* :601-603 (probsold = probs) */
        if ((!theta || !have0) && g_inner > 1 && !faithful_cost && S <= 32) {
            /* Same arithmetic with ONE pass over D per step and threads over independent outputs (see g_inner):
             * effect rows as bit masks, eff[i][s] bit j = closure_i[s, j]; a cell's row is the OR over its S-genes. */
            uint32_t *effrow = (uint32_t *)calloc((size_t)k * S, sizeof(uint32_t));
            for (int i = 0; i < k; i++)
                for (int s2 = 0; s2 < S; s2++)
                    for (int j = 0; j < S; j++)
                        if (clo[(size_t)i * SS + s2 + j * S]) effrow[i * S + s2] |= 1u << j;
            if (!theta) {
                /* :619 for all components: Wall[i][j][e] += d[e] * gamma_i[c] over ascending cells */
                double *Wall = (double *)calloc((size_t)k * E * (S + 1), sizeof(double));
                int nt = g_inner;
                if (nt > (E + 7) / 8) nt = (E + 7) / 8;
                int t;
#ifdef _OPENMP
#pragma omp parallel for schedule(static, 1) num_threads(nt)
#endif
                for (t = 0; t < nt; t++) {
                    const int per = ((E + nt - 1) / nt + 7) / 8 * 8;
                    const int e0 = t * per, e1 = e0 + per < E ? e0 + per : E;
                    for (int c = 0; c < C && e0 < e1; c++) {
                        const double *d = D + (size_t)c * E;
                        const uint32_t pm = pert_mask(pert[c]);
                        for (int i = 0; i < k; i++) {
                            const double g = post[i + (size_t)c * k];
                            uint32_t em = 0;
                            for (uint32_t r = pm; r; r &= r - 1) em |= effrow[i * S + __builtin_ctz(r)];
                            for (; em; em &= em - 1) {
                                double *w = Wall + ((size_t)i * (S + 1) + __builtin_ctz(em)) * E;
                                for (int e = e0; e < e1; e++) w[e] += d[e] * g;
                            }
                        }
                    }
                }
                for (int i = 0; i < k; i++) {
                    orc_max_col_row(Wall + (size_t)i * (S + 1) * E, E, S + 1, th + (size_t)i * E);   /* values 1..S+1 */
                    if (adv_here && !(g_nullcomp && i == k - 1))
                        adv_init_theta(Wall + (size_t)i * (S + 1) * E, E, S, th + (size_t)i * E,
                                       g_adv->init_theta + ((size_t)count * k + i) * E, count, i);
                }
                free(Wall);
            } else {
                for (size_t q = 0; q < (size_t)k * E; q++) th[q] = theta[q] == 0 ? S + 1 : theta[q];     /* :621-622 */
            }
            /* :624,631 L[i,c] = colSums(data * t(adj1[, subtopo])) : long-double, ascending e; cells are independent */
            ORC_PAR_FOR
            for (int c = 0; c < C; c++) {
                const double *d = D + (size_t)c * E;
                const uint32_t pm = pert_mask(pert[c]);
                for (int i = 0; i < k; i++) {
                    uint32_t em = 0;
                    for (uint32_t r = pm; r; r &= r - 1) em |= effrow[i * S + __builtin_ctz(r)];
                    const int32_t *thi = th + (size_t)i * E;
                    acc_t sum = 0;
                    for (int e = 0; e < E; e++) {
                        const int t2 = thi[e];
                        if (t2 <= S && ((em >> (t2 - 1)) & 1u)) sum += d[e];
                    }
                    probs0[i + (size_t)c * k] = (double)sum;
                }
            }
            free(effrow);
            have0 = 1;
        } else if (!theta || faithful_cost || !have0) {
            for (int i = 0; i < k; i++) {
                const uint8_t *A = clo + (size_t)i * SS;
                int32_t *thi = th + (size_t)i * E;
                if (!theta) {
                    /* :619 maxCol_row( (data * gamma_i) %*% cbind(adj1, 0) ), ascending-cell double sums */
                    memset(Wtmp, 0, sizeof(double) * (size_t)E * (S + 1));
                    for (int c = 0; c < C; c++) {
                        double g = post[i + (size_t)c * k];
                        const double *d = D + (size_t)c * E;
                        const uint32_t pm = pert_mask(pert[c]);
                        for (int j = 0; j < S; j++)
                            if (cell_effect(A, S, pm, j)) {
                                double *w = Wtmp + (size_t)j * E;
                                for (int e = 0; e < E; e++) w[e] += d[e] * g;
                            }
                    }
                    orc_max_col_row(Wtmp, E, S + 1, thi);                         /* values 1..S+1 */
                    if (adv_here && !(g_nullcomp && i == k - 1))
                        adv_init_theta(Wtmp, E, S, thi, g_adv->init_theta + ((size_t)count * k + i) * E, count, i);
                } else {
                    for (int e = 0; e < E; e++) { int t = theta[(size_t)i * E + e]; thi[e] = t == 0 ? S + 1 : t; }  /* :621-622 */
                }
                /* :624,631 L[i,c] = colSums(data * t(adj1[, subtopo])) : long-double, ascending e */
                for (int c = 0; c < C; c++) {
                    const double *d = D + (size_t)c * E;
                    const uint32_t pm = pert_mask(pert[c]);
                    uint8_t eff[ORC_MAXS];
                    for (int j = 0; j < S; j++) eff[j] = (uint8_t)cell_effect(A, S, pm, j);
                    acc_t s = 0;
                    for (int e = 0; e < E; e++) {
                        int t = thi[e];
                        if (t <= S && eff[t - 1]) s += d[e];
                    }
                    probs0[i + (size_t)c * k] = (double)s;
                }
            }
            have0 = 1;
        }
        if (g_nullcomp) {                                                        /* :607-608, 633: tmp <- 0 */
            for (int c = 0; c < C; c++) probs0[(k - 1) + (size_t)c * k] = 0.0;
            for (int e = 0; e < E; e++) th[(size_t)(k - 1) * E + e] = S + 1;
        }
        if (rec_probs) rec_theta(1, count, th, (size_t)k * E, S, 1);
        double ll0 = orc_get_ll(probs0, k, C, mw, complete, logtype);            /* :643 */
        int better_ll = ll0 - llold > 0;
        if (adv_here) {
            g_adv->n_checked[ADV_INIT_BEST]++;
            const int want = g_adv->init_best[count] != 0;
            if (want != better_ll) {
                const double mg = fabs(ll0 - llold) / (fabs(llold) > 1e-300 ? fabs(llold) : 1e-300);
                if (isfinite(mg) && mg <= g_adv->tol_ll) { adv_forced(ADV_INIT_BEST, mg); better_ll = want; }
                else adv_hard(ADV_INIT_BEST, mg, "best probs inside the initial getProbs", count, 0, 0);
            }
        }
        if (rec_probs) { rec_w(6); rec_w(g_rec_start); rec_w(count); rec_w(better_ll); }
        if (better_ll) memcpy(best, probs0, sizeof(double) * kc);                /* :646-648 */
        memcpy(probs, probs0, sizeof(double) * kc);                              /* :649 */
        mw_update_a(probs, k, C, mw, affinity, complete, logtype, 0, scratch);   /* :653-657 */
        llold = ll0;                                                             /* :658 */
    }
    memcpy(L_out, best, sizeof(double) * kc);
    for (int i = 0; i < k; i++) mw_out[i] = mw[i];
    if (theta_used) for (size_t i = 0; i < (size_t)k * E; i++) theta_used[i] = th[i] == S + 1 ? 0 : th[i];
    free(probs); free(probs0); free(best); free(post); free(scratch); free(clo); free(th); free(Wtmp);
    return llold;
}

double orc_get_probs(const double *D, int E, int C, const int32_t *pert, int S, int k, const uint8_t *phi,
                     const int32_t *theta, const double *L_in, const double *mw_in, int complete, double logtype,
                     int faithful_cost, double *L_out, double *mw_out, int32_t *theta_used)
{
    return orc_get_probs_a(D, E, C, pert, S, k, phi, theta, L_in, mw_in, 0, complete, logtype, faithful_cost, L_out, mw_out,
                           theta_used);
}

/* ------------------------------------------------------------------------------------------------ */
/* EM for one start: mnem()::do_inits, R/mnems.r:1884-2165, init != NULL branch (explicit phi).          */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    int max_trace;          /* capacity of the trace arrays */
    int complete;           /* mnem(complete=) */
    double logtype;         /* mnem(logtype=) */
    int faithful_cost;      /* recompute redundant work like the reference (timing only) */
    int max_iter_guard;     /* hard stop for runaway loops (the reference has none when increase=TRUE) */
    int noise_mode;         /* parity-harness aid, 0 = faithful.  The test `probs$ll > probs0$ll` (R/mnems.r:2104) is
                             * routinely decided by the last bit of ll once the graphs have stopped changing; 1 / 2
                             * resolve such noise-level (<= 1e-12 relative) comparisons as FALSE / TRUE so that both
                             * legitimate outcomes can be enumerated */
    int affinity;           /* mnem(affinity =): 0 soft responsibilities, 1 hard assignments in every getAffinity call */
    const double *mw0;      /* mnem(mw =): initial mixture weights (R/mnems.r:1771-1773); NULL = rep(1, k)/k */
    int nullcomp;           /* mnem(nullcomp = TRUE): the LAST of the k components is the null component (R/mnems.r:1832-1834,
                             * 1894-1900): empty graph, theta 0, log ratios 0 (R/mnems_low.r:607-608), no M-step (:2012, 2076-2080) */
    const int32_t *theta0;  /* mnem(theta =): k*E fixed attachments (0..S) of THIS start, or NULL.  R/mnems.r:1890-1892 puts them
                             * into res1, so the first getProbs takes its theta-given path; :2007-2011, 2033 hand them to nem() as
                             * `subtopo`, which returns them unchanged (R/mnems.r:362-372; the llr score ignores subtopo, :439-454) */
} orc_em_opts;

typedef struct {
    double ll;              /* limits$ll = max(lls) */
    double best_ll;
    int n_iter;             /* EM loop iterations executed (count) */
    long n_scored;          /* candidate graphs scored in all greedy searches */
    long n_greedy_iter;
    double min_margin;
    long n_raw;             /* candidates generated before unique() */
    double min_accept_margin;  /* smallest relative |ll_new - ll_kept| or |trace_a - trace_b| behind an EM-level decision */
    long n_dec, n_dec_tight;   /* greedy decisions of all searches, and those closer than 1e-13 relative (see orc_nem_info) */
} orc_em_info;

/* outputs: adj_out k*S*S (reduced, diag 0), theta_out k*E, L_out k*C (bestprobs), mw_out k (bestmw),
 * lls/phievo/thetaevo/mwevo traces of length n_iter. */
void orc_em_start_advised(const double *D, int E, int C, const int32_t *pert, int S, int k, const uint8_t *init_phi,
                          const orc_em_opts *o, orc_advice *adv, uint8_t *adj_out, int32_t *theta_out, double *L_out,
                          double *mw_out, double *lls, double *phievo, double *thetaevo, double *mwevo, orc_em_info *info);
void orc_em_start(const double *D, int E, int C, const int32_t *pert, int S, int k, const uint8_t *init_phi,
                  const orc_em_opts *o, uint8_t *adj_out, int32_t *theta_out, double *L_out, double *mw_out,
                  double *lls, double *phievo, double *thetaevo, double *mwevo, orc_em_info *info)
{
    orc_em_start_advised(D, E, C, pert, S, k, init_phi, o, NULL, adj_out, theta_out, L_out, mw_out, lls, phievo, thetaevo,
                         mwevo, info);
}

/* adv != NULL: replay of a device fit, see orc_advice.  adv == NULL: the plain restatement. */
void orc_em_start_advised(const double *D, int E, int C, const int32_t *pert, int S, int k, const uint8_t *init_phi,
                          const orc_em_opts *o, orc_advice *adv, uint8_t *adj_out, int32_t *theta_out, double *L_out,
                          double *mw_out, double *lls, double *phievo, double *thetaevo, double *mwevo, orc_em_info *info)
{
    g_adv = adv;
    if (adv) {
        for (int q = 0; q < ADV_KINDS; q++) { adv->n_checked[q] = adv->n_forced[q] = adv->n_hard[q] = 0; adv->max_forced[q] = 0;
                                              adv->min_hard[q] = INFINITY; }
        adv->diverged = 0; adv->first_hard[0] = 0;
        g_adv_probs = 1;
    }
    g_rec_probs = g_rec != NULL;
    double *Wfin0 = adv ? (double *)malloc(sizeof(double) * (size_t)E * S) : NULL;
    double *Wfin1 = adv ? (double *)malloc(sizeof(double) * (size_t)E * S) : NULL;
    const int SS = S * S;
    size_t kc = (size_t)k * C;
    double logtype = o->logtype;
    int complete = o->complete;
    double *probs = (double *)malloc(sizeof(double) * kc);        /* probs$probs */
    double *probsold = (double *)malloc(sizeof(double) * kc);
    double *newp = (double *)malloc(sizeof(double) * kc);
    double *post = (double *)malloc(sizeof(double) * kc);
    double *scratch = (double *)malloc(sizeof(double) * kc);
    double *gam = (double *)malloc(sizeof(double) * C);
    double *Rk = (double *)malloc(sizeof(double) * (size_t)E * S);
    double *Rall = NULL;
    uint8_t *res1_adj = (uint8_t *)malloc((size_t)k * SS), *res_adj = (uint8_t *)malloc((size_t)k * SS);
    int32_t *res1_th = (int32_t *)malloc(sizeof(int32_t) * (size_t)k * E), *res_th = (int32_t *)malloc(sizeof(int32_t) * (size_t)k * E);
    uint8_t *bestadj = (uint8_t *)malloc((size_t)k * SS);
    int32_t *bestth = (int32_t *)malloc(sizeof(int32_t) * (size_t)k * E);
    int have_res1_theta = 0;
    double mw[256], bestmw[256], mwnew[256];
    for (int i = 0; i < k; i++) mw[i] = o->mw0 ? o->mw0[i] : 1.0 / k;             /* R/mnems.r:1771-1773 */
    memcpy(res1_adj, init_phi, (size_t)k * SS);                                   /* :1885-1893 */
    if (o->nullcomp) memset(res1_adj + (size_t)(k - 1) * SS, 0, SS);              /* :1894-1898 */
    g_nullcomp = o->nullcomp;
    memset(res1_th, 0, sizeof(int32_t) * (size_t)k * E);
    if (o->theta0) { memcpy(res1_th, o->theta0, sizeof(int32_t) * (size_t)k * E); have_res1_theta = 1; }   /* :1890-1892 */
    memset(probs, 0, sizeof(double) * kc);                                        /* :1905 */
    const int affinity = o->affinity;
    double probs_ll = orc_get_probs_a(D, E, C, pert, S, k, res1_adj, o->theta0 ? res1_th : NULL, probs, mw, affinity, complete,
                                      logtype, o->faithful_cost, newp, mwnew, NULL);   /* :1906-1911 */
    memcpy(probs, newp, sizeof(double) * kc);
    for (int i = 0; i < k; i++) mw[i] = mwnew[i];                                  /* :1913 */
    mw_update_a(probs, k, C, mw, affinity, complete, logtype, 1, scratch);        /* :1936-1940 */
    double ll = -INFINITY, llold = -INFINITY, bestll = -INFINITY;                 /* :1945-1948 */
    for (int i = 0; i < k; i++) bestmw[i] = mw[i];
    memcpy(probsold, probs, sizeof(double) * kc);
    double probsold_ll = probs_ll;
    double *bestprobs = (double *)malloc(sizeof(double) * kc);
    memcpy(bestprobs, probs, sizeof(double) * kc);
    memcpy(bestadj, res1_adj, (size_t)k * SS);
    memset(bestth, 0, sizeof(int32_t) * (size_t)k * E);
    int count = 0, time0 = 1, stop = 0;
    long nscored = 0, ngreedy = 0, nraw = 0, ndec = 0, ntight = 0;
    double minmargin = INFINITY, acceptmargin = INFINITY;
    /* :1955-1958 with converged=-Inf, increase=TRUE: (ll - llold > -Inf & |ll - llold| > 0) | time0, & !stop.
     * `count < max_iter` binds only to the !increase arm (operator precedence). */
    for (;;) {
        int go = (((ll - llold > -INFINITY) && fabs(ll - llold) > 0) || time0) && !stop;
        if (adv_on()) {
            g_adv->n_checked[ADV_BEST]++;
            const int want = count < adv->n_em;
            if (want != go) {
                /* `stop` is integer arithmetic on graphs both sides share; what can differ is |ll - llold| > 0 */
                const double mg = fabs(ll - llold) / (fabs(ll) > 1e-300 ? fabs(ll) : 1e-300);
                if (!time0 && isfinite(mg) && mg <= adv->tol_ll && !(want && stop)) { adv_forced(ADV_BEST, mg); go = want; }
                else adv_hard(ADV_BEST, mg, "EM loop length", count, adv->n_em, 0);
            }
        }
        if (!go) break;
        if (!time0) {
            int upd = ll - bestll > 0;
            if (adv_on()) {
                g_adv->n_checked[ADV_BEST]++;
                const int want = adv->best_at[count] != 0;
                if (want != upd) {
                    const double mg = fabs(ll - bestll) / (fabs(ll) > 1e-300 ? fabs(ll) : 1e-300);
                    if (isfinite(mg) && mg <= adv->tol_ll) { adv_forced(ADV_BEST, mg); upd = want; }
                    else adv_hard(ADV_BEST, mg, "best result update", count, 0, 0);
                }
            }
            if (upd && g_rec) { rec_w(7); rec_w(g_rec_start); rec_w(count); }
            if (upd) {                                                            /* :1960-1963 */
                bestll = ll; memcpy(bestadj, res1_adj, (size_t)k * SS);
                memcpy(bestth, res1_th, sizeof(int32_t) * (size_t)k * E);
                memcpy(bestprobs, probs, sizeof(double) * kc);
                for (int i = 0; i < k; i++) bestmw[i] = mw[i];
            }
            llold = ll;                                                           /* :1964-1965 */
            memcpy(probsold, probs, sizeof(double) * kc); probsold_ll = probs_ll;
        }
        orc_get_affinity(probs, k, C, mw, affinity, complete, logtype, post);     /* :1968-1973 */
        double edgechange = 0, thetachange = 0;
        const int one_pass = g_inner > 1 && !o->faithful_cost;      /* all k collapsed matrices from one pass over D */
        if (one_pass) {
            if (!Rall) Rall = (double *)malloc(sizeof(double) * (size_t)k * E * S);
            do_mean_multi(D, E, C, pert, S, k, post, (size_t)k, Rall);
        }
        for (int i = 0; i < k; i++) {                                             /* :1976-2086 */
            if (o->nullcomp && i == k - 1) {                                      /* :2012, 2076-2080: no M-step, adj * 0, subtopo * 0 */
                memset(res_adj + (size_t)i * SS, 0, SS);
                memset(res_th + (size_t)i * E, 0, sizeof(int32_t) * E);
                for (int q = 0; q < SS; q++) edgechange += abs(0 - (int)res1_adj[(size_t)i * SS + q]);
                if (have_res1_theta)
                    for (int e = 0; e < E; e++) thetachange += 0 != res1_th[(size_t)i * E + e];
                continue;
            }
            uint8_t a0[ORC_MAXS * ORC_MAXS], a1[ORC_MAXS * ORC_MAXS];
            int32_t *t0 = (int32_t *)malloc(sizeof(int32_t) * E), *t1 = (int32_t *)malloc(sizeof(int32_t) * E);
            orc_nem_info i0, i1;
            if (one_pass) memcpy(Rk, Rall + (size_t)i * E * S, sizeof(double) * (size_t)E * S);
            else {
                for (int c = 0; c < C; c++) gam[c] = post[i + (size_t)c * k];
                orc_do_mean(D, E, C, pert, S, gam, Rk);                           /* nem :101-107 */
            }
            const int adv_it = adv_on() && count < adv->n_em;                     /* this EM iteration is on the tape */
            if (adv_it) { const size_t q = ((size_t)count * k + i) * 2;
                          g_adv_rows = adv->move_rows + (size_t)adv->move_off[q] * S;
                          g_adv_nmoves = (int)(adv->move_off[q + 1] - adv->move_off[q]); g_adv_W = Wfin0; }
            else if (adv) adv->diverged = adv->diverged || count >= adv->n_em;
            g_rec_iter = count; g_rec_comp = i; g_rec_restart = 0;
            orc_nem_greedy(Rk, E, S, NULL, a0, t0, &i0);                          /* j == 1: start0*0 :2042 */
            g_rec_restart = 1;
            if (o->faithful_cost) orc_do_mean(D, E, C, pert, S, gam, Rk);
            if (adv_it) { const size_t q = ((size_t)count * k + i) * 2 + 1;
                          g_adv_rows = adv->move_rows + (size_t)adv->move_off[q] * S;
                          g_adv_nmoves = (int)(adv->move_off[q + 1] - adv->move_off[q]); g_adv_W = Wfin1; }
            orc_nem_greedy(Rk, E, S, res1_adj + (size_t)i * SS, a1, t1, &i1);     /* j == 2: start0 :2040 */
            g_adv_rows = NULL; g_adv_nmoves = 0; g_adv_W = NULL;
            nscored += i0.n_scored + i1.n_scored; ngreedy += i0.n_iter + i1.n_iter; nraw += i0.n_raw + i1.n_raw;
            ndec += i0.n_dec + i1.n_dec; ntight += i0.n_dec_tight + i1.n_dec_tight;
            if (i0.min_margin < minmargin) minmargin = i0.min_margin;
            if (i1.min_margin < minmargin) minmargin = i1.min_margin;
            int pick1 = i1.max_trace > i0.max_trace;                              /* which.max: tie -> j == 1 :2072-2073 */
            if (i1.max_trace != i0.max_trace) {
                double mg = fabs(i1.max_trace - i0.max_trace) / (fabs(i0.max_trace) > 1e-300 ? fabs(i0.max_trace) : 1e-300);
                if (mg < acceptmargin) acceptmargin = mg;
            }
            if (adv_it && adv_on()) {
                g_adv->n_checked[ADV_PICK]++;
                const int want = adv->pick[(size_t)count * k + i] != 0;
                if (want != pick1) {
                    const double mg = fabs(i1.max_trace - i0.max_trace) / (fabs(i0.max_trace) > 1e-300 ? fabs(i0.max_trace) : 1e-300);
                    if (mg <= adv->tol) { adv_forced(ADV_PICK, mg); pick1 = want; }
                    else adv_hard(ADV_PICK, mg, "restart pick", count, i, 0);
                }
            }
            if (g_rec) { rec_w(3); rec_w(g_rec_start); rec_w(count); rec_w(i); rec_w(pick1); }
            memcpy(res_adj + (size_t)i * SS, pick1 ? a1 : a0, SS);
            memcpy(res_th + (size_t)i * E, pick1 ? t1 : t0, sizeof(int32_t) * E);
            if (o->theta0) memcpy(res_th + (size_t)i * E, o->theta0 + (size_t)i * E, sizeof(int32_t) * E);   /* subtopo = thetaX */
            if (adv_it && adv_on() && !o->theta0)
                adv_theta(ADV_THETA, pick1 ? Wfin1 : Wfin0, E, S, res_th + (size_t)i * E,
                          adv->theta + ((size_t)count * k + i) * E, count, i);
            for (int q = 0; q < SS; q++)                                          /* :2080-2081 */
                edgechange += abs((int)res_adj[(size_t)i * SS + q] - (int)res1_adj[(size_t)i * SS + q]);
            if (have_res1_theta)                                                  /* :2082-2083 (NULL != x sums to 0) */
                for (int e = 0; e < E; e++) thetachange += res_th[(size_t)i * E + e] != res1_th[(size_t)i * E + e];
            free(t0); free(t1);
        }
        rec_theta(4, count, res_th, (size_t)k * E, S, 0);
        memcpy(res1_adj, res_adj, (size_t)k * SS);                                /* :2092 */
        memcpy(res1_th, res_th, sizeof(int32_t) * (size_t)k * E);
        have_res1_theta = 1;
        /* E-step :2095-2108 */
        double newll = orc_get_probs_a(D, E, C, pert, S, k, res1_adj, res1_th, probsold, mw, affinity, complete, logtype,
                                       o->faithful_cost, newp, mwnew, NULL);
        if (isfinite(newll) && isfinite(probsold_ll)) {
            double mg = fabs(newll - probsold_ll) / (fabs(probsold_ll) > 1e-300 ? fabs(probsold_ll) : 1e-300);
            if (mg < acceptmargin) acceptmargin = mg;
        }
        if (o->noise_mode && isfinite(newll) && isfinite(probsold_ll) &&
            fabs(newll - probsold_ll) <= 1e-12 * fabs(probsold_ll)) {
            if (o->noise_mode == 1) newll = probsold_ll;                       /* "not greater" */
            else if (newll <= probsold_ll) newll = nextafter(probsold_ll, INFINITY);   /* "greater by one ulp" */
        }
        int take_new = newll > probsold_ll;
        if (adv_on() && count < adv->n_em) {
            g_adv->n_checked[ADV_ACCEPT]++;
            const int want = adv->accept[count] != 0;
            if (want != take_new) {
                const double mg = fabs(newll - probsold_ll) / (fabs(probsold_ll) > 1e-300 ? fabs(probsold_ll) : 1e-300);
                if (isfinite(mg) && mg <= adv->tol_ll) { adv_forced(ADV_ACCEPT, mg); take_new = want; }
                else adv_hard(ADV_ACCEPT, mg, "accept of the new E-step", count, 0, 0);
            }
        }
        if (g_rec) { rec_w(5); rec_w(g_rec_start); rec_w(count); rec_w(take_new); rec_w(1); }
        if (take_new) { memcpy(probs, newp, sizeof(double) * kc); probs_ll = newll; }
        else { memcpy(probs, probsold, sizeof(double) * kc); probs_ll = probsold_ll; }
        ll = probs_ll;                                                            /* :2111, evopen = 0 */
        double mwold[256];
        for (int i = 0; i < k; i++) mwold[i] = mw[i];
        mw_update_a(probs, k, C, mw, affinity, complete, logtype, 0, scratch);    /* :2113-2121 */
        acc_t dm = 0;
        for (int i = 0; i < k; i++) dm += fabs(mw[i] - mwold[i]);
        if (count < o->max_trace) { lls[count] = ll; phievo[count] = edgechange; thetaevo[count] = thetachange;
                                    mwevo[count] = (double)dm; }
        count++;
        if (edgechange == 0 && thetachange == 0 && !time0) stop = 1;              /* :2138-2141 */
        time0 = 0;
        if (o->max_iter_guard > 0 && count >= o->max_iter_guard) break;
    }
    int upd_final = ll - bestll > 0;
    if (adv_on()) {
        g_adv->n_checked[ADV_BEST]++;
        const int want = count <= adv->n_em && adv->best_at[count] != 0;
        if (want != upd_final) {
            const double mg = fabs(ll - bestll) / (fabs(ll) > 1e-300 ? fabs(ll) : 1e-300);
            if (isfinite(mg) && mg <= adv->tol_ll) { adv_forced(ADV_BEST, mg); upd_final = want; }
            else adv_hard(ADV_BEST, mg, "final best result update", count, 0, 0);
        }
    }
    if (g_rec) { if (upd_final) { rec_w(7); rec_w(g_rec_start); rec_w(count); } rec_w(8); rec_w(g_rec_start); rec_w(count); }
    if (upd_final) {                                                              /* :2144-2147 */
        bestll = ll; memcpy(bestadj, res1_adj, (size_t)k * SS);
        memcpy(bestth, res1_th, sizeof(int32_t) * (size_t)k * E);
        memcpy(bestprobs, probs, sizeof(double) * kc);
        for (int i = 0; i < k; i++) bestmw[i] = mw[i];
    }
    memcpy(adj_out, bestadj, (size_t)k * SS);
    memcpy(theta_out, bestth, sizeof(int32_t) * (size_t)k * E);
    memcpy(L_out, bestprobs, sizeof(double) * kc);
    for (int i = 0; i < k; i++) mw_out[i] = bestmw[i];
    double mx = -INFINITY;
    int nt = count < o->max_trace ? count : o->max_trace;
    for (int i = 0; i < nt; i++) if (lls[i] > mx) mx = lls[i];
    info->ll = mx;                                                                /* limits$ll <- max(lls) :2157 */
    info->best_ll = bestll; info->n_iter = count; info->n_scored = nscored; info->n_greedy_iter = ngreedy;
    info->min_margin = minmargin; info->n_raw = nraw; info->min_accept_margin = acceptmargin;
    info->n_dec = ndec; info->n_dec_tight = ntight;
    free(probs); free(probsold); free(newp); free(post); free(scratch); free(gam); free(Rk); free(Rall);
    free(res1_adj); free(res_adj); free(res1_th); free(res_th); free(bestadj); free(bestth); free(bestprobs);
    free(Wfin0); free(Wfin1);
    g_adv = NULL; g_adv_probs = 0; g_nullcomp = 0;
}
int orc_advice_size(void) { return (int)sizeof(orc_advice); }

/* All starts (lapply / sfLapply over do_inits, R/mnems.r:2166-2170) and the winner (:2222-2230: `<=` so the
 * LAST maximal start wins).  Per-start outputs are laid out start-major.  nthreads > 1 runs starts in
 * parallel like snowfall `parallel = n` (results are independent of the thread count).
 * Returns the 0-based index of the winning start. */
int orc_mnem_em(const double *D, int E, int C, const int32_t *pert, int S, int k, int starts,
                const uint8_t *init_phi, const orc_em_opts *o, int nthreads, uint8_t *adj_out, int32_t *theta_out,
                double *L_out, double *mw_out, double *lls, double *phievo, double *thetaevo, double *mwevo,
                orc_em_info *infos)
{
    const size_t SS = (size_t)S * S;
    int s;
#ifdef _OPENMP
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
#endif
    for (s = 0; s < starts; s++)
        orc_em_start(D, E, C, pert, S, k, init_phi + (size_t)s * k * SS, o, adj_out + (size_t)s * k * SS,
                     theta_out + (size_t)s * k * E, L_out + (size_t)s * k * C, mw_out + (size_t)s * k,
                     lls + (size_t)s * o->max_trace, phievo + (size_t)s * o->max_trace,
                     thetaevo + (size_t)s * o->max_trace, mwevo + (size_t)s * o->max_trace, infos + s);
    int best = 0;
    for (s = 1; s < starts; s++) if (infos[best].ll <= infos[s].ll) best = s;
    return best;
}

/* Final mixture weights of the returned object: lambda0, R/mnems.r:2298-2303. */
void orc_final_mw_a(const double *L, int k, int C, const double *mw_best, int affinity, int complete, double logtype,
                    double *out);
void orc_final_mw(const double *L, int k, int C, const double *mw_best, int complete, double logtype, double *out)
{
    orc_final_mw_a(L, k, C, mw_best, 0, complete, logtype, out);
}
void orc_final_mw_a(const double *L, int k, int C, const double *mw_best, int affinity, int complete, double logtype,
                    double *out)
{
    for (int i = 0; i < k; i++) out[i] = mw_best[i];
    if (k == 1) { out[0] = 1.0; return; }
    double *scratch = (double *)malloc(sizeof(double) * (size_t)k * C);
    mw_update_a(L, k, C, out, affinity, complete, logtype, 0, scratch);
    free(scratch);
}

int orc_acc_bits(void) { return (int)(sizeof(acc_t) * 8); }
int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
