/* TEST INFRASTRUCTURE ONLY -- toy implementation of the R / Biostrings C API declared in Rinternals.h and
 * Biostrings_interface.h (see there).  Objects are malloc'ed records kept in one list and freed by rstub_free_all(). */
#include "rstub.h"
extern "C" {
#include "Biostrings_interface.h"     /* plain C declarations, as in Biostrings */
}

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

struct rstub_object {
    int type;
    long len;
    void* data;         /* doubles, ints, SEXP* or chars (NUL-terminated) */
    SEXP names;
    int s4;
    /* XStringSet payload: one pool + starts / widths */
    char* pool;
    int* starts;
    int* widths;
};

jmp_buf rstub_top_level;
static char g_error[2048];
static int g_protect = 0;
static std::vector<SEXP>* g_all = nullptr;
static rstub_object g_nil = {NILSXP, 0, nullptr, nullptr, 0, nullptr, nullptr, nullptr};
static rstub_object g_names_symbol = {NILSXP, 0, nullptr, nullptr, 0, nullptr, nullptr, nullptr};
SEXP R_NilValue = &g_nil;
SEXP R_NamesSymbol = &g_names_symbol;

static SEXP new_object(int type, long len, size_t bytes) {
    if (!g_all) g_all = new std::vector<SEXP>();
    SEXP x = (SEXP)std::calloc(1, sizeof(rstub_object));
    x->type = type;
    x->len = len;
    x->data = bytes ? std::calloc(1, bytes) : nullptr;
    x->names = R_NilValue;
    g_all->push_back(x);
    return x;
}

extern "C" {

const char* rstub_last_error(void) { return g_error; }
int rstub_protect_depth(void) { return g_protect; }
void rstub_reset_protect(void) { g_protect = 0; }

void rstub_free_all(void) {
    if (!g_all) return;
    for (SEXP x : *g_all) {
        std::free(x->data);
        std::free(x->pool);
        std::free(x->starts);
        std::free(x->widths);
        std::free(x);
    }
    delete g_all;
    g_all = nullptr;
    g_protect = 0;
}

SEXP Rf_allocVector(unsigned int type, R_xlen_t n) {
    size_t el = 0;
    switch (type) {
        case REALSXP: el = sizeof(double); break;
        case INTSXP: case LGLSXP: el = sizeof(int); break;
        case STRSXP: case VECSXP: el = sizeof(SEXP); break;
        default: Rf_error("rstub: allocVector of unsupported type %u", type);
    }
    SEXP x = new_object((int)type, (long)n, el * (size_t)(n > 0 ? n : 1));
    if (type == STRSXP || type == VECSXP) {
        for (R_xlen_t i = 0; i < n; ++i) ((SEXP*)x->data)[i] = R_NilValue;
    }
    return x;
}

SEXP Rf_protect(SEXP x) { ++g_protect; return x; }
void Rf_unprotect(int n) { g_protect -= n; }
int LENGTH(SEXP x) { return (int)x->len; }
R_xlen_t Rf_xlength(SEXP x) { return (R_xlen_t)x->len; }
double* REAL(SEXP x) { return (double*)x->data; }
int* INTEGER(SEXP x) { return (int*)x->data; }
int* LOGICAL(SEXP x) { return (int*)x->data; }
SEXP VECTOR_ELT(SEXP x, R_xlen_t i) { return ((SEXP*)x->data)[i]; }
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v) { ((SEXP*)x->data)[i] = v; return v; }
SEXP STRING_ELT(SEXP x, R_xlen_t i) { return ((SEXP*)x->data)[i]; }
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v) { ((SEXP*)x->data)[i] = v; }
SEXP Rf_mkChar(const char* s) {
    const size_t n = std::strlen(s);
    SEXP x = new_object(CHARSXP, (long)n, n + 1);
    std::memcpy(x->data, s, n);
    return x;
}
const char* CHAR(SEXP x) { return (const char*)x->data; }
SEXP Rf_getAttrib(SEXP x, SEXP name) { return name == R_NamesSymbol ? x->names : R_NilValue; }
SEXP Rf_setAttrib(SEXP x, SEXP name, SEXP value) { if (name == R_NamesSymbol) x->names = value; return value; }
int IS_S4_OBJECT(SEXP x) { return x->s4; }
int Rf_isNumeric(SEXP x) { return x->type == REALSXP || x->type == INTSXP; }
int Rf_isString(SEXP x) { return x->type == STRSXP; }
int Rf_isLogical(SEXP x) { return x->type == LGLSXP; }
int Rf_isInteger(SEXP x) { return x->type == INTSXP; }
double Rf_asReal(SEXP x) { return x->type == REALSXP ? REAL(x)[0] : (double)INTEGER(x)[0]; }
int Rf_asLogical(SEXP x) { return LOGICAL(x)[0]; }

void Rf_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    std::vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    longjmp(rstub_top_level, 1);       /* like R: no C++ destructor between here and the top level runs */
}

/* ---- Biostrings ---- */
XStringSet_holder hold_XStringSet(SEXP x) {
    XStringSet_holder h;
    h.length = (int)x->len;
    h.opaque = x;
    return h;
}
int get_length_from_XStringSet_holder(const XStringSet_holder* x) { return x->length; }
Chars_holder get_elt_from_XStringSet_holder(const XStringSet_holder* x, int i) {
    SEXP o = (SEXP)x->opaque;
    Chars_holder c;
    c.ptr = o->pool + o->starts[i];
    c.length = o->widths[i];
    return c;
}
char DNAdecode(char code) {
    switch (code) {
        case 1: return 'A'; case 2: return 'C'; case 4: return 'G'; case 8: return 'T';
        case 3: return 'M'; case 5: return 'R'; case 9: return 'W'; case 6: return 'S'; case 10: return 'Y'; case 12: return 'K';
        case 7: return 'V'; case 11: return 'H'; case 13: return 'D'; case 14: return 'B'; case 15: return 'N';
        case 16: return '-'; case 32: return '+'; case 64: return '.';
    }
    Rf_error("DNAdecode: invalid code %d", (int)code);
}

/* ---- builders ---- */
SEXP rstub_string_vector(const char* const* strings, int n) {
    SEXP x = Rf_allocVector(STRSXP, n);
    for (int i = 0; i < n; ++i) SET_STRING_ELT(x, i, Rf_mkChar(strings[i]));
    return x;
}

static char dna_encode(char c) {
    switch (c) {
        case 'A': return 1; case 'C': return 2; case 'G': return 4; case 'T': return 8;
        case 'M': return 3; case 'R': return 5; case 'W': return 9; case 'S': return 6; case 'Y': return 10; case 'K': return 12;
        case 'V': return 7; case 'H': return 11; case 'D': return 13; case 'B': return 14; case 'N': return 15;
        case '-': return 16; case '+': return 32; case '.': return 64;
    }
    return 15;
}

SEXP rstub_xstringset(const char* const* strings, int n, int dna_codes) {
    SEXP x = new_object(S4SXP, n, 0);
    x->s4 = 1;
    size_t total = 0;
    for (int i = 0; i < n; ++i) total += std::strlen(strings[i]);
    x->pool = (char*)std::calloc(1, total + 1);
    x->starts = (int*)std::calloc((size_t)(n > 0 ? n : 1), sizeof(int));
    x->widths = (int*)std::calloc((size_t)(n > 0 ? n : 1), sizeof(int));
    size_t at = 0;
    for (int i = 0; i < n; ++i) {
        const size_t w = std::strlen(strings[i]);
        x->starts[i] = (int)at;
        x->widths[i] = (int)w;
        for (size_t k = 0; k < w; ++k) x->pool[at + k] = dna_codes ? dna_encode(strings[i][k]) : strings[i][k];
        at += w;
    }
    return x;
}

SEXP rstub_named_reals(const double* v, const char* const* names, int n) {
    SEXP x = Rf_allocVector(REALSXP, n);
    for (int i = 0; i < n; ++i) REAL(x)[i] = v[i];
    if (names) Rf_setAttrib(x, R_NamesSymbol, rstub_string_vector(names, n));
    return x;
}
SEXP rstub_real(double v) { return rstub_named_reals(&v, nullptr, 1); }
SEXP rstub_integers(const int* v, int n) {
    SEXP x = Rf_allocVector(INTSXP, n);
    for (int i = 0; i < n; ++i) INTEGER(x)[i] = v[i];
    return x;
}
SEXP rstub_logical(int v) {
    SEXP x = Rf_allocVector(LGLSXP, 1);
    LOGICAL(x)[0] = v;
    return x;
}
SEXP rstub_string(const char* s) { return rstub_string_vector(&s, 1); }
SEXP rstub_list(int n) { return Rf_allocVector(VECSXP, n); }

}  // extern "C"
