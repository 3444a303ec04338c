// C entry points around the reference's own src/mm.cpp, compiled UNMODIFIED from /root/reference with the stub headers in
// oracle/rstub (see oracle/Makefile: target _ref/libmm_ref.so).  TEST INFRASTRUCTURE ONLY.  Each function marshals
// plain arrays into the stub's SEXPs exactly as R's .Call would present them (column-major double matrices with a dim
// attribute, 1-based IntegerVector indices) and calls the reference routine.
#include <RcppEigen.h>
#include <Rcpp.h>

#include <cstring>

SEXP eigenMapMatMult(const Eigen::Map<Eigen::MatrixXd> A, Eigen::Map<Eigen::MatrixXd> B);
SEXP transClose_W(Rcpp::NumericMatrix x);
SEXP transClose_Del(Rcpp::NumericMatrix x, Rcpp::IntegerVector u, Rcpp::IntegerVector v);
SEXP transClose_Ins(Rcpp::NumericMatrix x, Rcpp::IntegerVector u, Rcpp::IntegerVector v);
SEXP maxCol_row(Rcpp::NumericMatrix x);

static SEXP int1(int v)
{
    SEXP s = Rf_allocVector(INTSXP, 1);
    INTEGER(s)[0] = v;
    return s;
}

extern "C" {
// x: n x n column-major doubles, modified in place like the reference does (src/mm.cpp:19,29)
void mmref_trans_close_w(double *x, int n) { transClose_W(Rcpp::NumericMatrix(rstub_wrap_real_matrix(x, n, n))); rstub_release(); }
void mmref_trans_close_del(double *x, int n, int u, int v)
{
    transClose_Del(Rcpp::NumericMatrix(rstub_wrap_real_matrix(x, n, n)), Rcpp::IntegerVector(int1(u)), Rcpp::IntegerVector(int1(v)));
    rstub_release();
}
void mmref_trans_close_ins(double *x, int n, int u, int v)
{
    transClose_Ins(Rcpp::NumericMatrix(rstub_wrap_real_matrix(x, n, n)), Rcpp::IntegerVector(int1(u)), Rcpp::IntegerVector(int1(v)));
    rstub_release();
}
// out[i] is pre-set by the caller: rows without any entry > -Inf stay untouched (the reference leaves them uninitialised)
void mmref_max_col_row(double *x, int nr, int nc, int *out)
{
    SEXP ans = maxCol_row(Rcpp::NumericMatrix(rstub_wrap_real_matrix(x, nr, nc)));
    // the reference's result vector is calloc'ed by the stub, so "uninitialised" reads as 0 here
    std::memcpy(out, INTEGER(ans), sizeof(int) * (size_t)nr);
    rstub_release();
}
void mmref_matmul(double *A, double *B, int m, int p, int n, double *C)
{
    SEXP r = eigenMapMatMult(Eigen::Map<Eigen::MatrixXd>(A, m, p), Eigen::Map<Eigen::MatrixXd>(B, p, n));
    std::memcpy(C, REAL(r), sizeof(double) * (size_t)m * n);
    rstub_release();
}
}
