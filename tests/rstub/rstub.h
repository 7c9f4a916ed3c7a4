/* TEST INFRASTRUCTURE ONLY -- what the driver needs beyond the R API: building inputs, catching Rf_error, cleaning up. */
#ifndef SARLACC_RSTUB_H
#define SARLACC_RSTUB_H

#include <setjmp.h>

#include "Rinternals.h"

#ifdef __cplusplus
extern "C" {
#endif

extern jmp_buf rstub_top_level;            /* where Rf_error lands */
const char* rstub_last_error(void);
int rstub_protect_depth(void);
void rstub_reset_protect(void);            /* what R does after an error */
void rstub_free_all(void);                 /* the "garbage collector": frees every object */

SEXP rstub_string_vector(const char* const* strings, int n);              /* character vector */
SEXP rstub_xstringset(const char* const* strings, int n, int dna_codes);  /* S4 XStringSet; dna_codes: bytes are Biostrings DNA codes */
SEXP rstub_named_reals(const double* v, const char* const* names, int n); /* names may be NULL */
SEXP rstub_real(double v);
SEXP rstub_integers(const int* v, int n);
SEXP rstub_logical(int v);
SEXP rstub_string(const char* s);
SEXP rstub_list(int n);

#ifdef __cplusplus
}
#endif

#endif
