// A small FUNCTIONAL stand-in for <Rcpp.h> -- TEST INFRASTRUCTURE.  R and Rcpp cannot be installed in the build image,
// so rshim/mnemcu_shim.cpp (the Rcpp glue a maintainer adds to the reference's src/) is compiled and run against this
// header instead: it models exactly the part of the Rcpp / R C API the shim touches (NumericMatrix, NumericVector,
// IntegerVector, List with names, Nullable<T>, as<T>, the `_["name"] = value` idiom, stop(), external pointers with
// finalizers) on reference-counted heap objects.  It checks the shim's syntax, types and marshalling; it says nothing
// about Rcpp itself.
#pragma once
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

typedef ptrdiff_t R_xlen_t;
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif

struct rstub_obj;
typedef std::shared_ptr<rstub_obj> SEXP;
struct rstub_obj {
    enum Kind { NIL, REAL, INT, LIST, EXTPTR } kind = NIL;
    std::vector<double> real;
    std::vector<int> integer;
    std::vector<SEXP> items;
    std::vector<std::string> names;
    int nrow = -1, ncol = -1;             // -1: a plain vector
    void *ptr = nullptr;
    void (*finalizer)(SEXP) = nullptr;
};
inline SEXP rstub_nil() { static SEXP n = std::make_shared<rstub_obj>(); return n; }
#define R_NilValue (rstub_nil())
inline bool Rf_isNull(const SEXP &s) { return !s || s->kind == rstub_obj::NIL; }
#define PROTECT(x) (x)
#define UNPROTECT(n) ((void)0)
inline SEXP R_MakeExternalPtr(void *p, SEXP, SEXP)
{
    SEXP s = std::make_shared<rstub_obj>();
    s->kind = rstub_obj::EXTPTR; s->ptr = p;
    return s;
}
inline void *R_ExternalPtrAddr(const SEXP &s) { return s->ptr; }
inline void R_ClearExternalPtr(const SEXP &s) { s->ptr = nullptr; }
inline void R_RegisterCFinalizerEx(const SEXP &s, void (*f)(SEXP), int) { s->finalizer = f; }
inline void rstub_finalize(const SEXP &s) { if (s->finalizer) s->finalizer(s); }      // what R's gc would do

namespace Rcpp {

[[noreturn]] inline void stop(const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw std::runtime_error(buf);
}

class NumericVector {
protected:
    SEXP s_;
public:
    NumericVector() : s_(std::make_shared<rstub_obj>()) { s_->kind = rstub_obj::REAL; }
    explicit NumericVector(int n) : NumericVector() { s_->real.assign(n, 0.0); }
    NumericVector(const SEXP &s) : s_(s) { if (s->kind != rstub_obj::REAL) stop("not a numeric vector"); }
    template <typename It> NumericVector(It a, It b) : NumericVector() { s_->real.assign(a, b); }
    static NumericVector create(double x) { NumericVector v(1); v[0] = x; return v; }
    double *begin() { return s_->real.data(); }
    const double *begin() const { return s_->real.data(); }
    double *end() { return s_->real.data() + s_->real.size(); }
    const double *end() const { return s_->real.data() + s_->real.size(); }
    R_xlen_t size() const { return (R_xlen_t)s_->real.size(); }
    double &operator[](R_xlen_t i) { return s_->real[i]; }
    double operator[](R_xlen_t i) const { return s_->real[i]; }
    operator SEXP() const { return s_; }
};

class NumericMatrix : public NumericVector {
public:
    NumericMatrix(int nr, int nc) : NumericVector(nr * nc) { s_->nrow = nr; s_->ncol = nc; }
    NumericMatrix(const SEXP &s) : NumericVector(s) { if (s->nrow < 0) stop("not a matrix"); }
    int nrow() const { return s_->nrow; }
    int ncol() const { return s_->ncol; }
};

class IntegerVector {
    SEXP s_;
public:
    IntegerVector() : s_(std::make_shared<rstub_obj>()) { s_->kind = rstub_obj::INT; }
    explicit IntegerVector(int n) : IntegerVector() { s_->integer.assign(n, 0); }
    IntegerVector(const SEXP &s) : s_(s) { if (s->kind != rstub_obj::INT) stop("not an integer vector"); }
    template <typename It> IntegerVector(It a, It b) : IntegerVector() { s_->integer.assign(a, b); }
    int *begin() { return s_->integer.data(); }
    const int *begin() const { return s_->integer.data(); }
    int *end() { return s_->integer.data() + s_->integer.size(); }
    R_xlen_t size() const { return (R_xlen_t)s_->integer.size(); }
    int &operator[](R_xlen_t i) { return s_->integer[i]; }
    int operator[](R_xlen_t i) const { return s_->integer[i]; }
    operator SEXP() const { return s_; }
};

inline SEXP wrap(const SEXP &s) { return s; }
inline SEXP wrap(double x) { return NumericVector::create(x); }
inline SEXP wrap(int x) { IntegerVector v(1); v[0] = x; return v; }

struct NamedValue { std::string name; SEXP value; };
struct NameProxy {
    std::string name;
    template <typename T> NamedValue operator=(const T &v) const { return NamedValue{name, wrap(v)}; }
};
struct NameMaker { NameProxy operator[](const char *n) const { return NameProxy{n}; } };
static const NameMaker _ = NameMaker();

class List {
    SEXP s_;
public:
    class Proxy {
        SEXP list_;
        size_t i_;
    public:
        Proxy(SEXP l, size_t i) : list_(l), i_(i) {}
        operator SEXP() const { return list_->items[i_]; }
        template <typename T> Proxy &operator=(const T &v) { list_->items[i_] = wrap(v); return *this; }
    };
    List() : s_(std::make_shared<rstub_obj>()) { s_->kind = rstub_obj::LIST; }
    explicit List(int n) : List() { s_->items.assign(n, R_NilValue); s_->names.assign(n, ""); }
    List(const SEXP &s) : s_(s) { if (s->kind != rstub_obj::LIST) stop("not a list"); }
    List(const Proxy &p) : List((SEXP)p) {}
    R_xlen_t size() const { return (R_xlen_t)s_->items.size(); }
    Proxy operator[](int i) { return Proxy(s_, (size_t)i); }
    SEXP operator[](int i) const { return s_->items[i]; }
    SEXP operator[](const char *name) const
    {
        for (size_t i = 0; i < s_->names.size(); ++i) if (s_->names[i] == name) return s_->items[i];
        stop("no list element named '%s'", name);
    }
    bool containsElementNamed(const char *name) const
    {
        for (const std::string &n : s_->names) if (n == name) return true;
        return false;
    }
    static List create() { return List(); }
    template <typename... Rest> static List create(const NamedValue &first, const Rest &...rest)
    {
        List l = create(rest...);
        l.s_->items.insert(l.s_->items.begin(), first.value);
        l.s_->names.insert(l.s_->names.begin(), first.name);
        return l;
    }
    operator SEXP() const { return s_; }
};
inline SEXP wrap(const List::Proxy &p) { return (SEXP)p; }

template <typename T> class Nullable {
    SEXP s_;
public:
    Nullable() : s_(R_NilValue) {}
    Nullable(const SEXP &s) : s_(s) {}
    Nullable(const T &t) : s_((SEXP)t) {}
    bool isNull() const { return Rf_isNull(s_); }
    bool isNotNull() const { return !isNull(); }
    operator SEXP() const { return s_; }
    operator T() const { return T(s_); }
};

template <typename T> T as(const SEXP &s) { return T(s); }
template <typename T> T as(const List::Proxy &p) { return T((SEXP)p); }

}  // namespace Rcpp
