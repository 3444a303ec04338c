// Minimal stand-in for <Rcpp.h> / the R C API, just enough to compile the reference's UNMODIFIED src/mm.cpp with plain g++
// (no R, Rcpp or Eigen exists in this image).  TEST INFRASTRUCTURE ONLY: used by oracle/Makefile to build
// oracle/_ref/libmm_ref.so, the compiled-reference checker for src/mm.cpp:15-84.  An "SEXP" here is a small tagged
// array; only the accessors mm.cpp touches exist (REAL, INTEGER, Rf_getAttrib(x, R_DimSymbol), Rf_allocVector,
// R_alloc, PROTECT/UNPROTECT, Rcpp::NumericMatrix, Rcpp::IntegerVector, Rcpp::wrap).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <limits>
#include <vector>

using std::round;

#define INTSXP 13
#define REALSXP 14

struct rstub_sexprec {
    int type = 0;
    long n = 0;
    int dim[2] = {0, 0};
    void *data = nullptr;      // owned unless `borrowed`
    bool borrowed = false;
    rstub_sexprec *dimattr = nullptr;
};
typedef rstub_sexprec *SEXP;

// every object and R_alloc block created during one glue call; released by rstub_release()
inline std::vector<SEXP> &rstub_objects() { static std::vector<SEXP> v; return v; }
inline std::vector<void *> &rstub_blocks() { static std::vector<void *> v; return v; }
inline void rstub_release()
{
    for (SEXP s : rstub_objects()) { if (!s->borrowed) free(s->data); delete s; }
    for (void *p : rstub_blocks()) free(p);
    rstub_objects().clear();
    rstub_blocks().clear();
}

inline SEXP Rf_allocVector(int type, long n)
{
    SEXP s = new rstub_sexprec();
    s->type = type; s->n = n;
    s->data = calloc(n > 0 ? n : 1, type == INTSXP ? sizeof(int) : sizeof(double));
    rstub_objects().push_back(s);
    return s;
}
inline SEXP rstub_wrap_real_matrix(double *p, int nr, int nc)       // R-owned matrix: the data stay the caller's
{
    SEXP s = new rstub_sexprec();
    s->type = REALSXP; s->n = (long)nr * nc; s->dim[0] = nr; s->dim[1] = nc; s->data = p; s->borrowed = true;
    rstub_objects().push_back(s);
    return s;
}
static SEXP R_DimSymbol = nullptr;
static const double R_NegInf = -std::numeric_limits<double>::infinity();
inline double *REAL(SEXP s) { return static_cast<double *>(s->data); }
inline int *INTEGER(SEXP s) { return static_cast<int *>(s->data); }
inline SEXP Rf_getAttrib(SEXP x, SEXP)                               // only ever asked for the dim attribute
{
    if (!x->dimattr) {
        x->dimattr = Rf_allocVector(INTSXP, 2);
        INTEGER(x->dimattr)[0] = x->dim[0];
        INTEGER(x->dimattr)[1] = x->dim[1];
    }
    return x->dimattr;
}
inline char *R_alloc(size_t n, int size)
{
    void *p = calloc(n > 0 ? n : 1, size);
    rstub_blocks().push_back(p);
    return static_cast<char *>(p);
}
#define PROTECT(x) (x)
#define UNPROTECT(n) ((void)0)

namespace Rcpp {
class NumericMatrix {
    SEXP s_;
public:
    NumericMatrix(SEXP s) : s_(s) {}
    operator SEXP() const { return s_; }
};
class IntegerVector {
    SEXP s_;
public:
    IntegerVector(SEXP s) : s_(s) {}
    int operator[](int i) const { return INTEGER(s_)[i]; }
    operator SEXP() const { return s_; }
};
inline SEXP wrap(SEXP s) { return s; }
inline SEXP wrap(const NumericMatrix &m) { return m; }
}  // namespace Rcpp
