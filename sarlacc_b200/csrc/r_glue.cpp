/* R-side glue: the four hot-path `.Call` entry points of sarlacc (and, at the end of the file, umi_group and
 * cluster_umis_test), re-implemented as thin SEXP unpackers over the C ABI in include/sarlacc_b200.h.  Drop this file into the package's src/ in place of
 * adaptor_align.cpp, barcode_align.cpp, general_align.cpp and reference_align.{h,cpp}; src/init.cpp keeps its
 * registration table (src/init.cpp:9-35) unchanged, because the symbols, arities and return shapes are the
 * same.  NOT compiled in this repository (no R / Rcpp / Biostrings headers in the image) -- it depends only on
 * Rinternals.h, Biostrings_interface.h and sarlacc_b200.h.  See INTEGRATION.md.
 *
 * Threading: called on R's main thread; SEXPs are only touched here, before and after the library call.  The
 * library reports errors by return code, so no C++ exception or longjmp crosses CUDA resources; Rf_error() is
 * raised from this outermost frame only (what BEGIN_RCPP/END_RCPP did in the reference).
 */
#include <Rinternals.h>
extern "C" {
#include "Biostrings_interface.h"
}
#include "sarlacc_b200.h"

#include <cstring>
#include <vector>

namespace {

struct Reads {   /* views layout: one (pointer, length) pair per XStringSet element, src/adaptor_align.cpp:46-50 */
    std::vector<const uint8_t*> sp, qp;
    std::vector<int32_t> sl, ql;
    sarlacc_reads r;
};

void hold(SEXP readseq, SEXP readqual, Reads& R) {
    XStringSet_holder q = hold_XStringSet(readqual);
    const int nq = get_length_from_XStringSet_holder(&q);
    int ns;
    if (IS_S4_OBJECT(readseq)) {            /* DNAStringSet: Biostrings byte codes, src/DNA_input.cpp:64-75 */
        XStringSet_holder s = hold_XStringSet(readseq);
        ns = get_length_from_XStringSet_holder(&s);
        R.sp.resize(ns); R.sl.resize(ns);
        for (int i = 0; i < ns; ++i) {
            Chars_holder e = get_elt_from_XStringSet_holder(&s, i);
            R.sp[i] = reinterpret_cast<const uint8_t*>(e.ptr);
            R.sl[i] = e.length;
        }
        R.r.seq_encoding = SARLACC_SEQ_BIOSTRINGS;
    } else {                                 /* character vector, src/DNA_input.cpp:47-51 */
        ns = LENGTH(readseq);
        R.sp.resize(ns); R.sl.resize(ns);
        for (int i = 0; i < ns; ++i) {
            SEXP e = STRING_ELT(readseq, i);
            R.sp[i] = reinterpret_cast<const uint8_t*>(CHAR(e));
            R.sl[i] = LENGTH(e);
        }
        R.r.seq_encoding = SARLACC_SEQ_ASCII;
    }
    if (ns != nq) Rf_error("sequence and quality vectors should have the same length");   /* :23-25 */
    R.qp.resize(nq); R.ql.resize(nq);
    for (int i = 0; i < nq; ++i) {
        Chars_holder e = get_elt_from_XStringSet_holder(&q, i);
        R.qp[i] = reinterpret_cast<const uint8_t*>(e.ptr);
        R.ql[i] = e.length;
    }
    R.r.n = ns;
    R.r.seq = R.sp.data(); R.r.seq_len = R.sl.data();
    R.r.qual = R.qp.data(); R.r.qual_len = R.ql.data();
    R.r.seq_pool = R.r.qual_pool = NULL; R.r.seq_off = R.r.qual_off = NULL;
}

struct Enc {
    std::vector<const char*> names;
    sarlacc_encoding e;
    explicit Enc(SEXP encoding) {
        SEXP nm = Rf_getAttrib(encoding, R_NamesSymbol);
        e.n = LENGTH(encoding);
        e.err = REAL(encoding);
        e.names = NULL;
        if (nm != R_NilValue && LENGTH(nm) == e.n) {
            names.resize(e.n);
            for (int i = 0; i < e.n; ++i) names[i] = CHAR(STRING_ELT(nm, i));
            e.names = names.data();
        }
    }
};

double numeric_scalar(SEXP x, const char* what) {     /* src/utils.cpp:18-20 */
    if (!Rf_isNumeric(x) || LENGTH(x) != 1) Rf_error("%s should be a numeric scalar", what);
    return Rf_asReal(x);
}

const char* string_scalar(SEXP x, const char* what) { /* src/utils.cpp:26-31 */
    if (!Rf_isString(x) || LENGTH(x) != 1) Rf_error("%s should be a string", what);
    return CHAR(STRING_ELT(x, 0));
}

void check(int rc) {
    if (rc != 0) Rf_error("%s", sarlacc_last_error());
}

}  // namespace

extern "C" {

SEXP adaptor_align(SEXP readseq, SEXP readqual, SEXP encoding, SEXP gapopen, SEXP gapext, SEXP adaptor, SEXP sec_starts, SEXP sec_ends) {
    const char* ad = string_scalar(adaptor, "adaptor sequence");
    const double go = numeric_scalar(gapopen, "gap opening penalty"), ge = numeric_scalar(gapext, "gap extension penalty");
    Reads R; hold(readseq, readqual, R);
    Enc E(encoding);
    const int nsec = LENGTH(sec_starts);
    if (nsec != LENGTH(sec_ends)) Rf_error("section starts and ends should have the same length");
    const int n = (int)R.r.n;
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 5));
    SEXP score = Rf_allocVector(REALSXP, n); SET_VECTOR_ELT(out, 0, score);
    SEXP start = Rf_allocVector(INTSXP, n);  SET_VECTOR_ELT(out, 1, start);
    SEXP end = Rf_allocVector(INTSXP, n);    SET_VECTOR_ELT(out, 2, end);
    SEXP ss = Rf_allocVector(VECSXP, nsec);  SET_VECTOR_ELT(out, 3, ss);
    SEXP sw = Rf_allocVector(VECSXP, nsec);  SET_VECTOR_ELT(out, 4, sw);
    std::vector<int32_t> sst((size_t)nsec * n), swd((size_t)nsec * n);   /* [nsec][n], copied into one IntegerVector per section */
    check(sarlacc_adaptor_align(&R.r, &E.e, go, ge, ad, nsec, INTEGER(sec_starts), INTEGER(sec_ends),
                                REAL(score), INTEGER(start), INTEGER(end), sst.data(), swd.data()));
    for (int s = 0; s < nsec; ++s) {
        SEXP a = Rf_allocVector(INTSXP, n); SET_VECTOR_ELT(ss, s, a);
        SEXP b = Rf_allocVector(INTSXP, n); SET_VECTOR_ELT(sw, s, b);
        for (int i = 0; i < n; ++i) { INTEGER(a)[i] = sst[(size_t)s * n + i]; INTEGER(b)[i] = swd[(size_t)s * n + i]; }
    }
    UNPROTECT(1);
    return out;
}

static SEXP score_only(SEXP readseq, SEXP readqual, SEXP encoding, SEXP gapopen, SEXP gapext, SEXP ref, const char* what, bool global) {
    const char* rf = string_scalar(ref, what);
    const double go = numeric_scalar(gapopen, "gap opening penalty"), ge = numeric_scalar(gapext, "gap extension penalty");
    Reads R; hold(readseq, readqual, R);
    Enc E(encoding);
    SEXP score = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)R.r.n));
    check(global ? sarlacc_barcode_align(&R.r, &E.e, go, ge, rf, REAL(score))
                 : sarlacc_adaptor_align_score_only(&R.r, &E.e, go, ge, rf, REAL(score)));
    UNPROTECT(1);
    return score;
}

SEXP adaptor_align_score_only(SEXP readseq, SEXP readqual, SEXP encoding, SEXP gapopen, SEXP gapext, SEXP adaptor) {
    return score_only(readseq, readqual, encoding, gapopen, gapext, adaptor, "adaptor sequence", false);
}

SEXP barcode_align(SEXP barcodeseq, SEXP barcodequal, SEXP encoding, SEXP gapopen, SEXP gapext, SEXP reference) {
    return score_only(barcodeseq, barcodequal, encoding, gapopen, gapext, reference, "barcode sequence", true);
}

SEXP general_align(SEXP inputseq, SEXP inputqual, SEXP encoding, SEXP gapopen, SEXP gapext, SEXP reference, SEXP edit_only) {
    const char* rf = string_scalar(reference, "reference sequence");
    const double go = numeric_scalar(gapopen, "gap opening penalty"), ge = numeric_scalar(gapext, "gap extension penalty");
    if (!Rf_isLogical(edit_only) || LENGTH(edit_only) != 1) Rf_error("edit-only specification should be a logical scalar");
    const int eo = Rf_asLogical(edit_only);
    Reads R; hold(inputseq, inputqual, R);
    Enc E(encoding);
    const int n = (int)R.r.n;
    int maxlen = 0;
    for (int i = 0; i < n; ++i) if (R.sl[i] > maxlen) maxlen = R.sl[i];
    const int64_t stride = (int64_t)maxlen + (int64_t)strlen(rf) + 2;
    std::vector<char> ra(eo ? 1 : (size_t)stride * n), qa(eo ? 1 : (size_t)stride * n);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 4));
    SEXP score = Rf_allocVector(REALSXP, n); SET_VECTOR_ELT(out, 0, score);
    SEXP edit = Rf_allocVector(INTSXP, n);   SET_VECTOR_ELT(out, 1, edit);
    SEXP rs = Rf_allocVector(STRSXP, eo ? 0 : n); SET_VECTOR_ELT(out, 2, rs);
    SEXP qs = Rf_allocVector(STRSXP, eo ? 0 : n); SET_VECTOR_ELT(out, 3, qs);
    check(sarlacc_general_align(&R.r, &E.e, go, ge, rf, eo, REAL(score), INTEGER(edit), ra.data(), qa.data(), stride));
    if (!eo) {
        for (int i = 0; i < n; ++i) {
            SET_STRING_ELT(rs, i, Rf_mkChar(ra.data() + (size_t)i * stride));
            SET_STRING_ELT(qs, i, Rf_mkChar(qa.data() + (size_t)i * stride));
        }
    }
    UNPROTECT(1);
    return out;
}

}

/* ---- UMI grouping: SEXP umi_group(umi1, thresh1, umi2, thresh2, pregroup) (src/umi_group.cpp:14-117, src/init.cpp:23)
 * and SEXP cluster_umis_test(links) (src/cluster_umis_test.cpp:8-29, src/init.cpp:25).  Replaces src/umi_group.cpp and
 * src/cluster_umis_test.cpp; R/umiGroup.R:21-22 keeps working unchanged because the result is nested per pre-group
 * exactly like the reference's (a list of lists of integer vectors, unlisted one level by the R code). */
namespace {

struct UmiPool {   /* decoded ASCII, what process_DNA_input()->get_persistent yields (src/DNA_input.cpp:28-88) */
    std::vector<uint8_t> pool;
    std::vector<int64_t> off;
    int64_t n = 0;
    explicit UmiPool(SEXP x) {
        off.push_back(0);
        if (IS_S4_OBJECT(x)) {
            XStringSet_holder h = hold_XStringSet(x);
            n = get_length_from_XStringSet_holder(&h);
            for (int64_t i = 0; i < n; ++i) {
                Chars_holder e = get_elt_from_XStringSet_holder(&h, (int)i);
                for (int k = 0; k < e.length; ++k) pool.push_back((uint8_t)DNAdecode(e.ptr[k]));
                off.push_back((int64_t)pool.size());
            }
        } else {
            n = LENGTH(x);
            for (int64_t i = 0; i < n; ++i) {
                SEXP e = STRING_ELT(x, i);
                pool.insert(pool.end(), CHAR(e), CHAR(e) + LENGTH(e));
                off.push_back((int64_t)pool.size());
            }
        }
        if (pool.empty()) pool.push_back(0);
    }
};

int integer_scalar(SEXP x, const char* what) {   /* check_integer_scalar, src/utils.cpp:14-16 */
    if (!Rf_isInteger(x) || LENGTH(x) != 1) Rf_error("%s should be an integer scalar", what);
    return INTEGER(x)[0];
}

/* sarlacc_lists -> VECSXP of INTSXP; frees the handle */
SEXP lists_to_sexp(sarlacc_lists* h, int64_t from, int64_t to, const std::vector<int64_t>& off, const std::vector<int32_t>& val) {
    (void)h;
    SEXP out = PROTECT(Rf_allocVector(VECSXP, (R_xlen_t)(to - from)));
    for (int64_t c = from; c < to; ++c) {
        SEXP v = Rf_allocVector(INTSXP, (R_xlen_t)(off[c + 1] - off[c]));
        SET_VECTOR_ELT(out, (R_xlen_t)(c - from), v);
        if (off[c + 1] > off[c]) std::memcpy(INTEGER(v), val.data() + off[c], sizeof(int32_t) * (size_t)(off[c + 1] - off[c]));
    }
    UNPROTECT(1);
    return out;
}

void fetch_lists(sarlacc_lists* h, std::vector<int64_t>& off, std::vector<int32_t>& val) {
    if (!h) Rf_error("%s", sarlacc_last_error());
    off.assign((size_t)sarlacc_lists_count(h) + 1, 0);
    val.assign((size_t)sarlacc_lists_values(h) + 1, 0);
    sarlacc_lists_fetch(h, off.data(), val.data());
    sarlacc_lists_free(h);
}

}

extern "C" {

SEXP umi_group(SEXP umi1, SEXP thresh1, SEXP umi2, SEXP thresh2, SEXP pregroup) {
    UmiPool u1(umi1);
    const int t1 = integer_scalar(thresh1, "threshold 1");
    const bool two = umi2 != R_NilValue;
    UmiPool u2(two ? umi2 : umi1);
    if (two && u1.n != u2.n) Rf_error("'umi1' and 'umi2' should have the same length");   /* src/umi_group.cpp:25-29 */
    const int t2 = integer_scalar(thresh2, "threshold 2");
    const R_xlen_t ng = Rf_xlength(pregroup);
    /* one library call per pre-group keeps the reference's nesting (a list per group) without a second index */
    SEXP out = PROTECT(Rf_allocVector(VECSXP, ng));
    for (R_xlen_t g = 0; g < ng; ++g) {
        SEXP cur = VECTOR_ELT(pregroup, g);
        const int64_t goff[2] = {0, (int64_t)LENGTH(cur)};
        std::vector<int64_t> off;
        std::vector<int32_t> val;
        fetch_lists(sarlacc_umi_group(u1.pool.data(), u1.off.data(), u1.n, t1, two ? u2.pool.data() : NULL, u2.off.data(), t2,
                                      goff, INTEGER(cur), 1, /*device*/ 0), off, val);
        SET_VECTOR_ELT(out, g, lists_to_sexp(NULL, 0, (int64_t)off.size() - 1, off, val));
    }
    UNPROTECT(1);
    return out;
}
/* (With many small pre-groups, pass them all in ONE call -- group_off / members as CSR -- and re-nest by counting the
 *  clusters per group: the device pass then covers every group at once.  sarlacc_b200/native.py: umi_group does that.) */

SEXP cluster_umis_test(SEXP links) {
    const R_xlen_t n = Rf_xlength(links);
    std::vector<int64_t> loff(1, 0);
    std::vector<int32_t> lval;
    for (R_xlen_t i = 0; i < n; ++i) {
        SEXP cur = VECTOR_ELT(links, i);
        lval.insert(lval.end(), INTEGER(cur), INTEGER(cur) + LENGTH(cur));
        loff.push_back((int64_t)lval.size());
    }
    if (lval.empty()) lval.push_back(0);
    std::vector<int64_t> off;
    std::vector<int32_t> val;
    fetch_lists(sarlacc_cluster_umis(loff.data(), lval.data(), (int64_t)n), off, val);
    return lists_to_sexp(NULL, 0, (int64_t)off.size() - 1, off, val);
}

}
