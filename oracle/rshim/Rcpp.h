/* TEST INFRASTRUCTURE ONLY -- not part of the product.
 *
 * A toy R object model + the slice of the Rcpp API that the reference's UMI-grouping sources touch, so that
 *   src/umi_group.cpp, src/sorted_trie.cpp, src/cluster_umis.cpp, src/cluster_umis_test.cpp, src/DNA_input.cpp,
 *   src/utils.cpp
 * compile VERBATIM where they lie under /root/reference (oracle/Makefile target `umiref`) and can be called through
 * their own extern "C" entry points (umi_group, fast_levdist_test, cluster_umis_test) by oracle/umi_ref_driver.cpp.
 * Objects live in an arena that the driver clears after each call; SEXP is a plain pointer into it.
 */
#ifndef SARLACC_ORACLE_RSHIM_RCPP_H
#define SARLACC_ORACLE_RSHIM_RCPP_H

#include <algorithm>
#include <cstddef>
#include <cstdarg>
#include <cstdio>
#include <deque>
#include <limits>
#include <memory>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

enum { RSHIM_NIL = 0, RSHIM_INT, RSHIM_REAL, RSHIM_LGL, RSHIM_STR, RSHIM_LIST, RSHIM_CHAR };

struct RshimObject {
    int type = RSHIM_NIL;
    bool s4 = false;
    std::vector<int> ints;                 /* INTSXP / LGLSXP */
    std::vector<double> reals;             /* REALSXP */
    std::vector<std::string> strs;         /* STRSXP */
    std::vector<RshimObject*> items;       /* VECSXP */
    std::string chars;                     /* CHARSXP */
};
typedef RshimObject* SEXP;

std::vector<std::unique_ptr<RshimObject> >& rshim_arena();      /* defined in the driver */
std::string& rshim_last_error();
inline SEXP rshim_new(int type) {
    rshim_arena().emplace_back(new RshimObject());
    rshim_arena().back()->type = type;
    return rshim_arena().back().get();
}
extern RshimObject rshim_nil_object;
#define R_NilValue (&rshim_nil_object)
#define R_NegInf (-std::numeric_limits<double>::infinity())

inline void Rprintf(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    std::vfprintf(stderr, fmt, ap);
    va_end(ap);
}

inline int Rf_length(SEXP x) {
    switch (x->type) {
        case RSHIM_INT: case RSHIM_LGL: return (int)x->ints.size();
        case RSHIM_REAL: return (int)x->reals.size();
        case RSHIM_STR: return (int)x->strs.size();
        case RSHIM_LIST: return (int)x->items.size();
        case RSHIM_CHAR: return (int)x->chars.size();
        default: return 0;
    }
}

#define BEGIN_RCPP try {
#define END_RCPP } catch (std::exception& rshim_e) { rshim_last_error() = rshim_e.what(); return nullptr; }

namespace Rcpp {

class RObject {
public:
    RObject(SEXP x = R_NilValue) : p(x) {}
    bool isS4() const { return p->s4; }
    operator SEXP() const { return p; }
    SEXP get() const { return p; }
private:
    SEXP p;
};

class String {
public:
    String() : p(rshim_new(RSHIM_CHAR)) {}
    String(const std::string& s) : p(rshim_new(RSHIM_CHAR)) { p->chars = s; }
    String(const char* s) : p(rshim_new(RSHIM_CHAR)) { p->chars = s; }
    const char* get_cstring() const { return p->chars.c_str(); }
    SEXP get_sexp() const { return p; }
    operator std::string() const { return p->chars; }
private:
    SEXP p;
};

template <class T>
inline T as(const std::string& s) { return T(s); }

class StringVector {
public:
    StringVector(RObject x) : p(x.get()) {
        if (p->type != RSHIM_STR) throw std::runtime_error("expecting a string vector");
    }
    explicit StringVector(size_t n) : p(rshim_new(RSHIM_STR)) { p->strs.resize(n); }
    size_t size() const { return p->strs.size(); }
    std::string& operator[](size_t i) { return p->strs[i]; }
    const std::string& operator[](size_t i) const { return p->strs[i]; }
    operator SEXP() const { return p; }
    operator RObject() const { return RObject(p); }
private:
    SEXP p;
};

/* Integer-like vectors share the object they were built from, like Rcpp handles do. */
template <int TYPE>
class IntLikeVector {
public:
    IntLikeVector() : p(rshim_new(TYPE)) {}
    IntLikeVector(SEXP x) : p(x) { check(); }
    IntLikeVector(RObject x) : p(x.get()) { check(); }
    IntLikeVector(size_t n, int value) : p(rshim_new(TYPE)) { p->ints.assign(n, value); }
    explicit IntLikeVector(size_t n) : p(rshim_new(TYPE)) { p->ints.assign(n, 0); }
    explicit IntLikeVector(int n) : p(rshim_new(TYPE)) { p->ints.assign((size_t)n, 0); }
    template <class It>
    IntLikeVector(It first, It last) : p(rshim_new(TYPE)) { p->ints.assign(first, last); }
    size_t size() const { return p->ints.size(); }
    int& operator[](size_t i) { return p->ints[i]; }
    const int& operator[](size_t i) const { return p->ints[i]; }
    std::vector<int>::iterator begin() { return p->ints.begin(); }
    std::vector<int>::iterator end() { return p->ints.end(); }
    std::vector<int>::const_iterator begin() const { return p->ints.begin(); }
    std::vector<int>::const_iterator end() const { return p->ints.end(); }
    operator SEXP() const { return p; }
    SEXP get() const { return p; }
private:
    void check() const {
        if (p->type != TYPE) throw std::runtime_error(TYPE == RSHIM_INT ? "expecting an integer vector" : "expecting a logical vector");
    }
    SEXP p;
};
typedef IntLikeVector<RSHIM_INT> IntegerVector;
typedef IntLikeVector<RSHIM_LGL> LogicalVector;

class NumericVector {
public:
    NumericVector(RObject x) : p(x.get()) {
        if (p->type != RSHIM_REAL) throw std::runtime_error("expecting a numeric vector");
    }
    size_t size() const { return p->reals.size(); }
    double& operator[](size_t i) { return p->reals[i]; }
private:
    SEXP p;
};

class List {
public:
    class Proxy {
    public:
        Proxy(SEXP owner, size_t i) : o(owner), idx(i) {}
        operator IntegerVector() const { return IntegerVector(o->items[idx]); }
        operator SEXP() const { return o->items[idx]; }
        Proxy& operator=(const IntegerVector& v) { o->items[idx] = v.get(); return *this; }
        Proxy& operator=(const List& v) { o->items[idx] = v.get(); return *this; }
    private:
        SEXP o;
        size_t idx;
    };
    List() : p(rshim_new(RSHIM_LIST)) {}
    List(SEXP x) : p(x) {
        if (p->type != RSHIM_LIST) throw std::runtime_error("expecting a list");
    }
    explicit List(size_t n) : p(rshim_new(RSHIM_LIST)) { p->items.assign(n, R_NilValue); }
    template <class It>
    List(It first, It last) : p(rshim_new(RSHIM_LIST)) {
        for (; first != last; ++first) p->items.push_back(SEXP(*first));
    }
    static List create(const IntegerVector& v) {
        List out(static_cast<size_t>(1));
        out.p->items[0] = v.get();
        return out;
    }
    size_t size() const { return p->items.size(); }
    Proxy operator[](size_t i) { return Proxy(p, i); }
    operator SEXP() const { return p; }
    SEXP get() const { return p; }
private:
    SEXP p;
};

}

#endif
