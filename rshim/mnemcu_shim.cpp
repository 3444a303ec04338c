// Rcpp pass-through onto libmnemcu.so.  R, Rcpp and their headers are not installable in the build image, so here the
// file is compiled and EXECUTED against a small functional stand-in for <Rcpp.h> (tests/cabi/rcpp_stub/Rcpp.h,
// driver tests/cabi/shim_check.cpp, run by tests/test_gpu_parity.py) -- that checks syntax, types and the marshalling,
// not Rcpp itself.  It is deliberately a thin marshalling layer: every line of numerical logic lives behind the C ABI (include/mnemcu.h) and is exercised
// through the same symbols by the Python host mirror (mnem_b200/) and its tests.
//
// Drop-in: add this file and include/mnemcu.h to the reference's src/, link with -lmnemcu (src/Makevars:
// PKG_LIBS = -L$(MNEMCU_HOME) -lmnemcu), run Rcpp::compileAttributes().  The five legacy exports keep their names
// and signatures (src/mm.cpp:8-84, R/RcppExports.R:4-22) so mytc(), llrScore(), getProbs() and scoreAdj() work
// unchanged; the coarse-grained entries below are what mnem()/nem() dispatch to when options(mnem.backend="cuda").
//
// [[Rcpp::depends(Rcpp)]]
#include <Rcpp.h>

#include "mnemcu.h"

using namespace Rcpp;

static void chk(int rc)
{
    if (rc != 0) stop("libmnemcu: %s", mnemcu_last_error());   // C status -> R error, like BEGIN_RCPP/END_RCPP
}

// graphs are double 0/1 matrices in R (closure kernels round(), src/mm.cpp:24-25); the ABI takes uint8
static std::vector<uint8_t> to_u8(const NumericMatrix &x)
{
    std::vector<uint8_t> g(x.size());
    for (R_xlen_t i = 0; i < x.size(); ++i) g[i] = std::round(x[i]) != 0;
    return g;
}
static void from_u8(const std::vector<uint8_t> &g, NumericMatrix &x)
{
    for (R_xlen_t i = 0; i < x.size(); ++i) x[i] = g[i];
}

// ---- legacy symbols (same names as src/mm.cpp) -----------------------------------------------------------------
// [[Rcpp::export]]
SEXP eigenMapMatMult(NumericMatrix A, NumericMatrix B)
{
    NumericMatrix C(A.nrow(), B.ncol());
    chk(mnemcu_matmul(A.begin(), B.begin(), A.nrow(), A.ncol(), B.ncol(), C.begin()));
    return C;
}

// [[Rcpp::export]]
SEXP transClose_W(NumericMatrix x)          // mutates x in place like the reference (src/mm.cpp:19,29)
{
    std::vector<uint8_t> g = to_u8(x);
    chk(mnemcu_trans_close_w(g.data(), x.nrow(), 1));
    from_u8(g, x);
    return x;
}

// [[Rcpp::export]]
SEXP transClose_Del(NumericMatrix x, IntegerVector u, IntegerVector v)
{
    std::vector<uint8_t> g = to_u8(x);
    int32_t uu = u[0], vv = v[0];
    chk(mnemcu_trans_close_del(g.data(), x.nrow(), 1, &uu, &vv));
    from_u8(g, x);
    return x;
}

// [[Rcpp::export]]
SEXP transClose_Ins(NumericMatrix x, IntegerVector u, IntegerVector v)
{
    std::vector<uint8_t> g = to_u8(x);
    int32_t uu = u[0], vv = v[0];
    chk(mnemcu_trans_close_ins(g.data(), x.nrow(), 1, &uu, &vv));
    from_u8(g, x);
    return x;
}

// [[Rcpp::export]]
SEXP maxCol_row(NumericMatrix x)
{
    IntegerVector out(x.nrow());
    chk(mnemcu_max_col_row(x.begin(), x.nrow(), x.ncol(), out.begin()));
    return out;
}

// ---- device-resident data: an external pointer with a finalizer ------------------------------------------------
static void ctx_finalizer(SEXP p)
{
    mnemcu_ctx *cx = static_cast<mnemcu_ctx *>(R_ExternalPtrAddr(p));
    if (cx) mnemcu_ctx_destroy(cx);
    R_ClearExternalPtr(p);
}

// data: E x C after modData(); pert = as.integer(colnames(data))  (R/mnems_low.r:415-433)
// [[Rcpp::export]]
SEXP mnemcu_context(NumericMatrix data, IntegerVector pert, int S, int device = 0)
{
    mnemcu_ctx *cx = nullptr;
    chk(mnemcu_ctx_create(data.begin(), data.nrow(), data.ncol(), pert.begin(), S, device, &cx));
    SEXP p = PROTECT(R_MakeExternalPtr(cx, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(p, ctx_finalizer, TRUE);
    UNPROTECT(1);
    return p;
}
// Rho = getRho(data) (R/mnems_low.r:308-324): S x C, several knock-downs per column allowed
// [[Rcpp::export]]
SEXP mnemcu_context_rho(NumericMatrix data, NumericMatrix Rho, int device = 0)
{
    if (Rho.ncol() != data.ncol()) stop("Rho must have one column per column of data");
    mnemcu_ctx *cx = nullptr;
    chk(mnemcu_ctx_create_rho(data.begin(), data.nrow(), data.ncol(), Rho.begin(), Rho.nrow(), device, &cx));
    SEXP p = PROTECT(R_MakeExternalPtr(cx, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(p, ctx_finalizer, TRUE);
    UNPROTECT(1);
    return p;
}
static mnemcu_ctx *ctx_of(SEXP p)
{
    mnemcu_ctx *cx = static_cast<mnemcu_ctx *>(R_ExternalPtrAddr(p));
    if (!cx) stop("mnemcu context was released");
    return cx;
}

// ---- getAffinity / getLL (R/mnems.r:1390, R/mnems_low.r:540) ---------------------------------------------------
// [[Rcpp::export]]
SEXP mnemcu_getAffinity(NumericMatrix x, NumericVector mw, int affinity, bool complete, double logtype)
{
    NumericMatrix y(x.nrow(), x.ncol());
    chk(mnemcu_get_affinity(x.begin(), x.nrow(), x.ncol(), mw.begin(), affinity, complete, logtype, y.begin()));
    return y;
}

// [[Rcpp::export]]
double mnemcu_getLL(NumericMatrix x, NumericVector mw, bool complete, double logtype)
{
    double ll = 0;
    chk(mnemcu_get_ll(x.begin(), x.nrow(), x.ncol(), mw.begin(), complete, logtype, &ll));
    return ll;
}

// ---- doMean / scoreAdj / nem greedy (R/mnems_low.r:871, R/mnems.r:416, :203-307) -------------------------------
// [[Rcpp::export]]
SEXP mnemcu_doMean(SEXP ctx, Nullable<NumericVector> weights)
{
    mnemcu_ctx *cx = ctx_of(ctx);
    int E, C, S;
    chk(mnemcu_ctx_dims(cx, &E, &C, &S));
    NumericMatrix R(E, S);
    chk(mnemcu_do_mean(cx, weights.isNull() ? nullptr : NumericVector(weights).begin(), 1, R.begin()));
    return R;
}

// [[Rcpp::export]]
List mnemcu_scoreAdj(NumericMatrix R, NumericMatrix adj, bool trans_close, bool marginal = false, double logtype = 2.0)
{
    const int E = R.nrow(), S = R.ncol();
    std::vector<uint8_t> g = to_u8(adj);
    double score = 0;
    IntegerVector theta(E);
    NumericMatrix W(E, S);
    if (marginal)
        chk(mnemcu_score_adj_marginal(R.begin(), E, S, g.data(), 1, trans_close, logtype, &score, theta.begin(), W.begin()));
    else
        chk(mnemcu_score_adj(R.begin(), E, S, g.data(), 1, trans_close, &score, theta.begin(), W.begin()));
    return List::create(_["score"] = score, _["subtopo"] = theta, _["subweights"] = W);
}

// [[Rcpp::export]]
List mnemcu_nem(NumericMatrix R, Nullable<NumericMatrix> start, bool marginal = false, double logtype = 2.0)
{
    const int E = R.nrow(), S = R.ncol();
    std::vector<uint8_t> st, adj((size_t)S * S);
    if (start.isNotNull()) st = to_u8(NumericMatrix(start));
    IntegerVector theta(E);
    double score = 0, mtr = 0;
    int32_t nit = 0;
    int64_t nsc = 0;
    if (marginal)
        chk(mnemcu_nem_greedy_marginal(R.begin(), E, S, start.isNotNull() ? st.data() : nullptr, 1, logtype, adj.data(),
                                       theta.begin(), &score, &mtr, &nit, &nsc));
    else
        chk(mnemcu_nem_greedy(R.begin(), E, S, start.isNotNull() ? st.data() : nullptr, 1, adj.data(), theta.begin(),
                              &score, &mtr, &nit, &nsc));
    NumericMatrix A(S, S);
    from_u8(adj, A);
    return List::create(_["adj"] = A, _["score"] = score, _["scores"] = NumericVector::create(mtr), _["subtopo"] = theta);
}

// ---- getProbs (R/mnems_low.r:576) -------------------------------------------------------------------------------
// res: list of k lists with $adj (S x S) and optionally $subtopo (all or none)
// [[Rcpp::export]]
List mnemcu_getProbs(SEXP ctx, NumericMatrix probs, List res, NumericVector mw, bool complete, double logtype)
{
    mnemcu_ctx *cx = ctx_of(ctx);
    int E, C, S;
    chk(mnemcu_ctx_dims(cx, &E, &C, &S));
    const int k = res.size();
    std::vector<uint8_t> phi((size_t)k * S * S);
    std::vector<int32_t> theta;
    bool have_theta = true;
    for (int i = 0; i < k; ++i) {
        List ri = res[i];
        std::vector<uint8_t> g = to_u8(as<NumericMatrix>(ri["adj"]));
        std::copy(g.begin(), g.end(), phi.begin() + (size_t)i * S * S);
        if (!ri.containsElementNamed("subtopo") || Rf_isNull(ri["subtopo"])) have_theta = false;
    }
    if (have_theta) {
        theta.resize((size_t)k * E);
        for (int i = 0; i < k; ++i) {
            IntegerVector t = as<IntegerVector>(as<List>(res[i])["subtopo"]);
            std::copy(t.begin(), t.end(), theta.begin() + (size_t)i * E);
        }
    }
    NumericMatrix L(k, C);
    NumericVector mwo(k);
    double ll = 0;
    chk(mnemcu_get_probs(cx, 1, k, phi.data(), have_theta ? theta.data() : nullptr, probs.begin(), mw.begin(), complete,
                         logtype, L.begin(), mwo.begin(), &ll, nullptr));
    return List::create(_["probs"] = L, _["subtopoX"] = R_NilValue, _["mw"] = mwo, _["ll"] = ll);
}

// ---- the EM phase of mnem(D, phi = init) (R/mnems.r:1884-2165, 2222-2230) ---------------------------------------
// init: list over starts of lists over components of S x S matrices, exactly mnem()'s `phi` argument
// [[Rcpp::export]]
// affinity: mnem(affinity =) 0 / 1 (R/mnems.r:1393-1417); mw: mnem(mw =) or NULL for rep(1, k)/k (:1771-1773);
// nullcomp: mnem(nullcomp = TRUE) -- init then carries k + 1 graphs per start, the last one the null component
// (:1832-1834, 1894-1898); theta_fixed: mnem(theta =), list over starts of lists over components of integer vectors, or NULL
List mnemcu_em(SEXP ctx, List init, bool complete, double logtype, int batch = 0, int max_trace = 256, int affinity = 0,
               Nullable<NumericVector> mw = R_NilValue, bool nullcomp = false, Nullable<List> theta_fixed = R_NilValue)
{
    mnemcu_ctx *cx = ctx_of(ctx);
    int E, C, S;
    chk(mnemcu_ctx_dims(cx, &E, &C, &S));
    const int starts = init.size(), k = as<List>(init[0]).size();
    std::vector<uint8_t> phi((size_t)starts * k * S * S), adj(phi.size());
    for (int s = 0; s < starts; ++s)
        for (int i = 0; i < k; ++i) {
            std::vector<uint8_t> g = to_u8(as<NumericMatrix>(as<List>(init[s])[i]));
            std::copy(g.begin(), g.end(), phi.begin() + ((size_t)s * k + i) * S * S);
        }
    std::vector<int32_t> theta((size_t)starts * k * E), nit(starts);
    std::vector<double> probs((size_t)starts * k * C), mwout((size_t)starts * k), ll(starts);
    std::vector<double> lls((size_t)starts * max_trace), pe(lls.size()), te(lls.size()), me(lls.size());
    int32_t best = -1;
    std::vector<double> mw0;
    if (mw.isNotNull()) {
        NumericVector m(mw);
        if ((int)m.size() != k) stop("mw must hold one weight per component");
        mw0.assign(m.begin(), m.end());
    }
    std::vector<int32_t> th0;
    if (theta_fixed.isNotNull()) {
        List tl(theta_fixed);
        if ((int)tl.size() != starts) stop("theta must hold one list per start");
        th0.resize((size_t)starts * k * E);
        for (int s = 0; s < starts; ++s)
            for (int i = 0; i < k; ++i) {
                IntegerVector t = as<IntegerVector>(as<List>(tl[s])[i]);
                if ((int)t.size() != E) stop("theta vectors must have one entry per E-gene");
                std::copy(t.begin(), t.end(), th0.begin() + ((size_t)s * k + i) * E);
            }
    }
    mnemcu_em_opts o = {complete, logtype, batch, max_trace, 0, nullptr, nullptr, affinity, nullptr, 0, 0,
                        mw0.empty() ? nullptr : mw0.data(), nullcomp ? 1 : 0, th0.empty() ? nullptr : th0.data()};
    mnemcu_em_stats st;
    chk(mnemcu_em_run(cx, starts, k, phi.data(), &o, adj.data(), theta.data(), probs.data(), mwout.data(), ll.data(),
                      nit.data(), lls.data(), pe.data(), te.data(), me.data(), &best, &st));
    List limits(starts);
    for (int s = 0; s < starts; ++s) {
        List res(k);
        for (int i = 0; i < k; ++i) {
            NumericMatrix A(S, S);
            std::copy(adj.begin() + ((size_t)s * k + i) * S * S, adj.begin() + ((size_t)s * k + i + 1) * S * S, A.begin());
            IntegerVector t(theta.begin() + ((size_t)s * k + i) * E, theta.begin() + ((size_t)s * k + i + 1) * E);
            res[i] = List::create(_["adj"] = A, _["subtopo"] = t);
        }
        NumericMatrix P(k, C);
        std::copy(probs.begin() + (size_t)s * k * C, probs.begin() + (size_t)(s + 1) * k * C, P.begin());
        const int n = std::min<int>(nit[s], max_trace);
        limits[s] = List::create(
            _["probs"] = List::create(_["probs"] = P), _["res"] = res, _["ll"] = ll[s],
            _["lls"] = NumericVector(lls.begin() + (size_t)s * max_trace, lls.begin() + (size_t)s * max_trace + n),
            _["mw"] = NumericVector(mwout.begin() + (size_t)s * k, mwout.begin() + (size_t)(s + 1) * k),
            _["phievo"] = NumericVector(pe.begin() + (size_t)s * max_trace, pe.begin() + (size_t)s * max_trace + n),
            _["thetaevo"] = NumericVector(te.begin() + (size_t)s * max_trace, te.begin() + (size_t)s * max_trace + n),
            _["mwevo"] = NumericVector(me.begin() + (size_t)s * max_trace, me.begin() + (size_t)s * max_trace + n));
    }
    return List::create(_["limits"] = limits, _["best"] = best + 1);
}
