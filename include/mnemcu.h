/*
 * mnemcu.h -- C ABI of libmnemcu.so, the sm_100a CUDA implementation of M&NEM's EM hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch/R types.  The Rcpp shim in rshim/
 * (source only; R is not installable in this image) and the Python host mirror mnem_b200/ bind exactly
 * these symbols.  Every entry point names the reference interface it replaces (paths relative to the
 * reference checkout, cbg-ethz/mnem 1.23.0).
 *
 * Conventions
 *   - all matrices column-major, FP64, exactly as R stores them (a cell's E values are contiguous);
 *   - graphs are uint8 0/1, S x S column-major, adj[i + j*S] = 1 <=> edge i -> j; a batch of n graphs is n
 *     consecutive S*S blocks; S <= 32 in this build;
 *   - theta (E-gene attachment, `subtopo`) is int32[E], values 0..S, 0 = unattached (R/mnems.r:448-449);
 *   - pert is int32[C], values 1..S: the S-gene knocked down in each cell, i.e. as.integer(colnames(D))
 *     after modData() (R/mnems_low.r:415-433).  Data with multi-knock-down columns ("1_3") is given by its
 *     S x C perturbation matrix Rho = getRho(D) (R/mnems_low.r:308-324) through mnemcu_ctx_create_rho;
 *   - `probs`/L is the k x C matrix of per-cell log ratios in units of log_b (R/mnems.r:1674-1675);
 *   - host pointers in, host pointers out; the caller owns every buffer.  Only mnemcu_ctx keeps device
 *     memory between calls; no host pointer is retained after a call returns;
 *   - every function returns 0 on success, non-zero otherwise with a thread-local message available from
 *     mnemcu_last_error().  No exceptions or longjmp cross this boundary.  There is no CPU fallback: without
 *     a CUDA device every compute entry point fails with MNEMCU_ERR_CUDA.
 */
#ifndef MNEMCU_H
#define MNEMCU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MNEMCU_OK 0
#define MNEMCU_ERR_ARG 1
#define MNEMCU_ERR_CUDA 2
#define MNEMCU_ERR_NOMEM 3
#define MNEMCU_MAX_S 32
#define MNEMCU_MAX_K 32

typedef struct mnemcu_ctx mnemcu_ctx;

/* ---- library ----------------------------------------------------------------------------------------- */
int mnemcu_version(void);
const char *mnemcu_last_error(void);
int mnemcu_device_count(int *n);
/* number of kernels this library has launched in this process (bench.py reports it as gpu_launches) */
int64_t mnemcu_launch_count(void);

/* ---- the five native routines of src/mm.cpp (registered in src/RcppExports.cpp:14-87) -------------------- */
/* transClose_W  src/mm.cpp:15-30    Warshall closure in place, n graphs.  (mytc() sets the diagonal first.) */
int mnemcu_trans_close_w(uint8_t *x, int S, int n);
/* transClose_Ins src/mm.cpp:50-65   x[i,j] |= x[i,u] && x[v,j]; u[g], v[g] 1-based per graph. */
int mnemcu_trans_close_ins(uint8_t *x, int S, int n, const int32_t *u, const int32_t *v);
/* transClose_Del src/mm.cpp:32-48   x[u,v] |= exists k: x[u,k] && x[k,v]. */
int mnemcu_trans_close_del(uint8_t *x, int S, int n, const int32_t *u, const int32_t *v);
/* maxCol_row    src/mm.cpp:66-84    per row first strict maximum, 1-based (0 for an all-NaN/-Inf row). */
int mnemcu_max_col_row(const double *x, int nr, int nc, int32_t *out);
/* eigenMapMatMult src/mm.cpp:8-13   C(m x n) = A(m x p) * B(p x n), ascending-p FP64 sums. */
int mnemcu_matmul(const double *A, const double *B, int m, int p, int n, double *C);

/* ---- graph helpers ----------------------------------------------------------------------------------- */
/* transitive.reduction R/mnems.r:511-533 (closure, zero diagonal, order-dependent in-place pruning). */
int mnemcu_trans_reduce(const uint8_t *x, int S, int n, uint8_t *out);
/* c(get.ins.fast, get.rev.tc, get.del.tc)(P) R/mnems_low.r:1025-1084, in that order, WITHOUT the unique() of
 * R/mnems.r:222 (duplicates have equal scores and which.max keeps the first, so results do not depend on it).
 * models: capacity 3*S*S graphs; kind[i] 0 ins / 1 rev / 2 del; *n_out = number written. */
int mnemcu_neighbours(const uint8_t *P, int S, uint8_t *models, int8_t *kind, int32_t *n_out);

/* ---- responsibilities and likelihood ----------------------------------------------------------------- */
/* getAffinity R/mnems.r:1390-1441 (norm = TRUE).  mw may be NULL (uniform).  affinity 0 soft / 1 hard. */
int mnemcu_get_affinity(const double *L, int k, int C, const double *mw, int affinity, int complete,
                        double logtype, double *out);
/* getLL R/mnems_low.r:540-552 */
int mnemcu_get_ll(const double *L, int k, int C, const double *mw, int complete, double logtype, double *ll);

/* ---- structure search on a collapsed E x S matrix ---------------------------------------------------- */
/* scoreAdj(D = R, adj, trans.close, dotopo = TRUE) R/mnems.r:416-458 for n candidate graphs against one R.
 * score[n]; theta (n*E) and W (n*E*S, `subweights`) may be NULL. */
int mnemcu_score_adj(const double *R, int E, int S, const uint8_t *adj, int n, int trans_close, double *score,
                     int32_t *theta, double *W);
/* scoreAdj(..., marginal = TRUE, logtype) R/mnems.r:455-457: score = sum_e log_b(sum_j b^cbind(0, R %*% adj)[e, j]);
 * theta and W as above (the attachment is the same argmax). */
int mnemcu_score_adj_marginal(const double *R, int E, int S, const uint8_t *adj, int n, int trans_close,
                              double logtype, double *score, int32_t *theta, double *W);
/* nem(search = "greedy", runs = 1) R/mnems.r:111-142,203-307,362-379 for n independent searches:
 * R is n blocks of E x S, start n graphs (NULL = empty starts).  Outputs per search: adj (transitively
 * reduced), theta, score (nem$score), max_trace (max(nem$scores)), n_iter accepted moves, n_scored
 * candidate graphs scored (before unique()). */
int mnemcu_nem_greedy(const double *R, int E, int S, const uint8_t *start, int n, uint8_t *adj_out,
                      int32_t *theta_out, double *score, double *max_trace, int32_t *n_iter, int64_t *n_scored);
/* nem(search = "greedy", marginal = TRUE, logtype): the same hill-climb with every scoreAdj call in its marginal
 * form (R/mnems.r:135, 208, 363). */
int mnemcu_nem_greedy_marginal(const double *R, int E, int S, const uint8_t *start, int n, double logtype,
                               uint8_t *adj_out, int32_t *theta_out, double *score, double *max_trace,
                               int32_t *n_iter, int64_t *n_scored);

/* ---- device-resident data context -------------------------------------------------------------------- */
/* Uploads D (E x C) once, sorted by perturbation, in the two HBM layouts the kernels stream. */
int mnemcu_ctx_create(const double *D, int E, int C, const int32_t *pert, int S, int device, mnemcu_ctx **out);
/* Same as mnemcu_ctx_create, but D is ALREADY in the memory of `device` (column-major E x C, the caller's cell
 * order; the buffer is only read and may be freed after the call).  For N-GPU jobs: every rank uploads 1/N of the
 * columns and the ranks all-gather over NVLink instead of pushing N full copies through PCIe (bench.py e2e). */
int mnemcu_ctx_create_device(const double *D_dev, int E, int C, const int32_t *pert, int S, int device,
                             mnemcu_ctx **out);
/* The Rho path of the reference (mnem(), nem(), getProbs(), doMean() called with Rho != NULL): Rho is S x C
 * column-major, Rho[s,c] != 0 <=> S-gene s is knocked down in cell c (any number per cell, also none).  Every
 * entry point that takes the context then follows the reference's Rho branches: effect of a cell =
 * min(1, t(Rho) %*% closure(phi)) (R/mnems_low.r:613-616), doMean = (D diag w) %*% t(Rho') with the columns
 * of Rho' normalised to sum 1 (:896-904).  The collapsed E x S matrix handed to the structure search is an
 * ordinary single-knock-down problem (nem() re-derives Rho from it, R/mnems.r:101-107). */
int mnemcu_ctx_create_rho(const double *D, int E, int C, const double *Rho, int S, int device, mnemcu_ctx **out);
/* Same, but D is produced on the device by the simData-style generator (mnem_b200/simdata.py describes the
 * tables): comp[C] component of each cell (0-based), phi_true n_comp graphs (closed), theta_true n_comp*E
 * (0 = uninformative row).  D[e,c] = 2*bin - 1 + noise(seed, e + E*c). */
int mnemcu_ctx_create_synthetic(int E, int C, const int32_t *pert, int S, const uint8_t *comp, int n_comp,
                                const uint8_t *phi_true, const int32_t *theta_true, uint64_t seed, int device,
                                mnemcu_ctx **out);
/* Host-side twin of the generator (fills cells [c0, c0+nc) of D, column-major E x nc); bit-identical. */
int mnemcu_synthetic_fill_host(double *D, int E, int64_t c0, int64_t nc, const int32_t *pert, int S,
                               const uint8_t *comp, int n_comp, const uint8_t *phi_true,
                               const int32_t *theta_true, uint64_t seed, int nthreads);
int mnemcu_ctx_destroy(mnemcu_ctx *ctx);
int mnemcu_ctx_dims(const mnemcu_ctx *ctx, int *E, int *C, int *S);
/* copy cells [c0, c0+nc) of the resident D back to the host (column-major E x nc) */
int mnemcu_ctx_read_data(mnemcu_ctx *ctx, int64_t c0, int64_t nc, double *D_out);

/* doMean(D, weights, Rho) R/mnems_low.r:871-920 on the Rho path taken from nem() (R/mnems.r:75-77,101-107):
 * R[j][e,s] = sum_{c: pert(c) = s} D[e,c] * weights[j][c].  weights: n x C row blocks (NULL: n = 1, no weights).
 * R_out: n blocks of E x S. */
int mnemcu_do_mean(mnemcu_ctx *ctx, const double *weights, int n, double *R_out);

/* getProbs R/mnems_low.r:576-663 for n independent mixtures of k components (n*k graphs, theta n*k*E or NULL
 * for the argmax path :618-619, L_in n blocks of k x C, mw_in n*k or NULL).  Outputs: L_out n blocks k x C
 * (`probs`), mw_out n*k, ll_out n, theta_used n*k*E (may be NULL).  The inner loop runs its full 10 (C > 100)
 * or 100 iterations as it does when called from mnem() with converged = -Inf. */
int mnemcu_get_probs(mnemcu_ctx *ctx, int n, int k, const uint8_t *phi, const int32_t *theta, const double *L_in,
                     const double *mw_in, int complete, double logtype, double *L_out, double *mw_out,
                     double *ll_out, int32_t *theta_used);

/* ---- the EM loop ------------------------------------------------------------------------------------- */
typedef struct {
    int complete;        /* mnem(complete =) */
    double logtype;      /* mnem(logtype =), default 2 */
    int batch;           /* EM starts sharing one pass over D (0 = choose so that k*batch <= 20) */
    int max_trace;       /* capacity per start of the lls/phievo/thetaevo/mwevo arrays */
    int max_iter_guard;  /* hard stop per start, 0 = none (the reference has none when increase = TRUE) */
    /* Optional shared work queue (several callers, one per GPU, fitting one set of starts): when non-NULL the
     * next start index comes from next_start(user) -- any value outside 0..starts-1 means "drained" -- instead
     * of 0, 1, 2, ...  Starts this caller did not fit keep n_iter = -1 and ll = -Inf.  The reference's analogue
     * is the snowfall worker pool that do_inits is mapped over (R/mnems.r:2166-2170). */
    int32_t (*next_start)(void *user);
    void *next_start_user;
    int affinity;        /* mnem(affinity =): 0 soft responsibilities (default), 1 hard assignments in every
                          * getAffinity call of the loop (R/mnems.r:1393-1417); not with complete = TRUE */
    /* Optional decision tape for audits (NULL = off; costs a few device-to-host copies per EM iteration): every
     * decision of the fit that compares floating-point sums, as int32 records [code, start, index, ...]:
     *   1 start inner  theta[k*E]             theta inside the theta-free getProbs (R/mnems_low.r:618-619), inner iteration
     *   6 start inner  flag                    `ll0 - llold > 0` there (:646-648)
     *   2 start iter comp restart n rows[n*S]  a greedy search (R/mnems.r:203-291): the graph (S row masks, bit j of row s =
     *                                          edge s -> j, diagonal set) after each of its n accepted moves
     *   3 start iter comp pick                 which restart was kept (:2069-2073)
     *   4 start iter   theta[k*E]              theta after the M-step of EM iteration iter
     *   5 start iter   accept improved         `probs$ll > probs0$ll` (:2104) and whether getProbs improved on its input
     *   7 start count                          best result replaced after `count` EM iterations (:1960-1963, :2144-2147)
     *   8 start n_iter                         end of the start
     * tests/ replay it through the CPU oracle (oracle/mnem_oracle.c, orc_advice). */
    int32_t *tape;
    int64_t tape_cap;    /* capacity of `tape` in int32 words; a fit that needs more fails and reports the need in
                          * mnemcu_em_stats::tape_words */
    int probs_mode;      /* layout of probs_out: 0 = one k x C block per start (starts * k * C doubles);
                          * 1 = ONE k x C block, that of the winning start among those this call fitted -- what mnem()
                          * returns (R/mnems.r:2222-2240) without keeping k x C doubles per start on the host */
    const double *mw;    /* mnem(mw =): k initial mixture weights of every start (R/mnems.r:1771-1773, handed to the first
                          * getProbs :1906-1913); NULL = rep(1, k)/k */
    int nullcomp;        /* mnem(nullcomp = TRUE): the LAST of the k components is the null component (R/mnems.r:1832-1834,
                          * 1894-1900): its graph is empty, its theta 0, its log ratios 0 (R/mnems_low.r:607-608) and it has
                          * no M-step (R/mnems.r:2012, 2076-2080); the last graph of every start in init_phi is ignored */
    const int32_t *theta; /* mnem(theta =): starts * k * E fixed E-gene attachments in 0..S (start-major), or NULL.  They go
                          * into res1 (R/mnems.r:1890-1892: the first getProbs takes its theta-given path) and are handed to
                          * every nem() call as `subtopo` (:2007-2011), which returns them unchanged: the graphs are still
                          * searched, theta never moves */
} mnemcu_em_opts;

typedef struct {
    int64_t em_iters;        /* EM loop iterations summed over starts (metric m1) */
    int64_t cand_scored;     /* candidate graphs scored in all greedy searches (metric m2) */
    int64_t greedy_iters;
    int64_t passes_estep;    /* launches of the E-step streaming kernel */
    int64_t passes_mstats;
    double ms_total, ms_estep, ms_mstats, ms_search, ms_mixture;   /* CUDA-event times */
    double estep_bytes;      /* algorithmic bytes moved by all E-step launches */
    double search_sm_busy;   /* device-resident greedy loop: time its CTAs spent on tasks / time they were resident */
    int64_t tape_words;      /* int32 words of the decision tape (0 when off) */
} mnemcu_em_stats;

/* mnem(D, phi = init) EM phase: do_inits for every start, R/mnems.r:1884-2165 (explicit-phi entry :1792-1793).
 * init_phi: starts*k graphs.  Per-start outputs (start-major): adj (k graphs, reduced), theta (k*E), probs
 * (k x C = limits$probs$probs), mw (k = limits$mw), ll (limits$ll), n_iter, traces (max_trace each).
 * *best = index of the winning start by R/mnems.r:2222-2230 (last maximal ll) among the starts fitted by this
 * call (-1 if it fitted none).  stats may be NULL. */
int mnemcu_em_run(mnemcu_ctx *ctx, int starts, int k, const uint8_t *init_phi, const mnemcu_em_opts *opts,
                  uint8_t *adj_out, int32_t *theta_out, double *probs_out, double *mw_out, double *ll_out,
                  int32_t *n_iter_out, double *lls, double *phievo, double *thetaevo, double *mwevo, int32_t *best,
                  mnemcu_em_stats *stats);

/* Timing probe for bench.py: n_rep launches of the E-step kernel alone for `batch` mixtures of k components
 * with the given graphs/theta (device-resident inputs), CUDA-event time per launch in ms_out[n_rep]. */
int mnemcu_bench_estep(mnemcu_ctx *ctx, int batch, int k, const uint8_t *phi, const int32_t *theta, int n_rep,
                       double *ms_out, double *bytes_per_launch);
int mnemcu_bench_mstats(mnemcu_ctx *ctx, int batch, int k, int n_rep, double *ms_out, double *bytes_per_launch);

#ifdef __cplusplus
}
#endif
#endif
