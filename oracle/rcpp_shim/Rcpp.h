/* TEST INFRASTRUCTURE ONLY -- not part of the product.
 *
 * Minimal stand-in for <Rcpp.h>, just large enough that the reference's own
 * src/reference_align.cpp and src/quality_encoding.cpp (compiled where they lie
 * under /root/reference, never copied) build without R.  It provides only the
 * handful of Rcpp names those two files touch:
 *   Rcpp::NumericVector  (size, [], names)         quality_encoding.cpp:5-32
 *   Rcpp::StringVector   (size, [])                quality_encoding.cpp:6-14
 *   Rcpp::as<std::string>                          quality_encoding.cpp:14
 *   R_NegInf, Rprintf                              reference_align.cpp:77,96,122
 */
#ifndef SARLACC_ORACLE_RCPP_SHIM_H
#define SARLACC_ORACLE_RCPP_SHIM_H

#include <cstddef>
#include <cstdio>
#include <cstdarg>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#define R_NegInf (-std::numeric_limits<double>::infinity())

inline void Rprintf(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    std::vfprintf(stderr, fmt, ap);
    va_end(ap);
}

namespace Rcpp {

class StringVector {
public:
    StringVector() : v(std::make_shared<std::vector<std::string> >()) {}
    explicit StringVector(const std::vector<std::string>& s) : v(std::make_shared<std::vector<std::string> >(s)) {}
    size_t size() const { return v->size(); }
    const std::string& operator[](size_t i) const { return (*v)[i]; }
private:
    std::shared_ptr<std::vector<std::string> > v;
};

template <class T>
inline T as(const std::string& s) { return T(s); }

/* Shares storage on copy, like an R vector handle. */
class NumericVector {
public:
    NumericVector() : d(std::make_shared<std::vector<double> >()) {}
    NumericVector(const double* x, size_t n, const std::vector<std::string>& nm) :
        d(std::make_shared<std::vector<double> >(x, x + n)), nms(nm) {}
    size_t size() const { return d->size(); }
    double& operator[](size_t i) { return (*d)[i]; }
    const double& operator[](size_t i) const { return (*d)[i]; }
    StringVector names() const { return nms; }
private:
    std::shared_ptr<std::vector<double> > d;
    StringVector nms;
};

}

#endif
