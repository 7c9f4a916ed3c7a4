/* TEST INFRASTRUCTURE ONLY -- a stand-in for R's C API, just large enough to compile and drive
 * sarlacc_b200/csrc/r_glue.cpp without an R installation (none exists in this image).
 *
 * The object model is a toy (tests/rstub/rstub.cpp), but the one property of the real API that the glue has to respect is
 * kept: Rf_error() does not return -- it longjmp()s to the caller's top level, skipping every C++ destructor on the way.
 * The driver (tests/rstub/glue_driver.cpp) runs each call under setjmp and under AddressSanitizer, so glue code that
 * raises an R error while std::vector or std::string objects are alive shows up as a leak.
 */
#ifndef SARLACC_RSTUB_RINTERNALS_H
#define SARLACC_RSTUB_RINTERNALS_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rstub_object* SEXP;
typedef ptrdiff_t R_xlen_t;

enum { NILSXP = 0, CHARSXP = 9, LGLSXP = 10, INTSXP = 13, REALSXP = 14, STRSXP = 16, VECSXP = 19, S4SXP = 25 };

extern SEXP R_NilValue;
extern SEXP R_NamesSymbol;

SEXP Rf_allocVector(unsigned int type, R_xlen_t n);
SEXP Rf_protect(SEXP x);
void Rf_unprotect(int n);
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)

int LENGTH(SEXP x);
R_xlen_t Rf_xlength(SEXP x);
double* REAL(SEXP x);
int* INTEGER(SEXP x);
int* LOGICAL(SEXP x);
SEXP VECTOR_ELT(SEXP x, R_xlen_t i);
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v);
SEXP STRING_ELT(SEXP x, R_xlen_t i);
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v);
SEXP Rf_mkChar(const char* s);
const char* CHAR(SEXP x);
SEXP Rf_getAttrib(SEXP x, SEXP name);
SEXP Rf_setAttrib(SEXP x, SEXP name, SEXP value);
int IS_S4_OBJECT(SEXP x);
int Rf_isNumeric(SEXP x);
int Rf_isString(SEXP x);
int Rf_isLogical(SEXP x);
int Rf_isInteger(SEXP x);
double Rf_asReal(SEXP x);
int Rf_asLogical(SEXP x);
#if defined(__GNUC__)
void Rf_error(const char* fmt, ...) __attribute__((noreturn, format(printf, 1, 2)));
#else
void Rf_error(const char* fmt, ...);
#endif

#ifdef __cplusplus
}
#endif

#endif
