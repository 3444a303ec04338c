// Minimal stand-in for <RcppEigen.h>: Eigen::MatrixXd, Eigen::Map<MatrixXd> and their product, enough for
// eigenMapMatMult (src/mm.cpp:8-13) to compile.  Eigen itself is a third-party dependency that is not vendored in
// /root/reference, so the product below is a RESTATEMENT of what Eigen's GEBP computes for these shapes (depth <= a few
// hundred, SSE2 build without FMA): one accumulator per output, ascending inner index, separate multiply and add.
// It pins nothing about Eigen; the four loop kernels of mm.cpp (:15-84) are the part that is compiled verbatim.
#pragma once
#include "Rcpp.h"

namespace Eigen {
class MatrixXd {
public:
    int r = 0, c = 0;
    std::vector<double> v;      // column-major
    MatrixXd() = default;
    MatrixXd(int rows, int cols) : r(rows), c(cols), v((size_t)rows * cols, 0.0) {}
    int rows() const { return r; }
    int cols() const { return c; }
    const double *data() const { return v.data(); }
};
template <class M>
class Map {
public:
    double *p;
    int r, c;
    Map(double *ptr, int rows, int cols) : p(ptr), r(rows), c(cols) {}
};
inline MatrixXd operator*(const Map<MatrixXd> &A, const Map<MatrixXd> &B)
{
    MatrixXd C(A.r, B.c);
    for (int j = 0; j < B.c; ++j)
        for (int i = 0; i < A.r; ++i) {
            double s = 0.0;
            for (int l = 0; l < A.c; ++l) s += A.p[i + (size_t)l * A.r] * B.p[l + (size_t)j * B.r];
            C.v[i + (size_t)j * A.r] = s;
        }
    return C;
}
}  // namespace Eigen

namespace Rcpp {
inline SEXP wrap(const Eigen::MatrixXd &m)
{
    SEXP s = Rf_allocVector(REALSXP, (long)m.rows() * m.cols());
    s->dim[0] = m.rows(); s->dim[1] = m.cols();
    for (long i = 0; i < s->n; ++i) REAL(s)[i] = m.data()[i];
    return s;
}
}  // namespace Rcpp
