/* R-side glue: the four hot-path `.Call` entry points of sarlacc (then two optional fused routines and, at the end of the
 * file, umi_group and cluster_umis_test), re-implemented as thin SEXP unpackers over the C ABI in include/sarlacc_b200.h.  Drop this file into
 * the package's src/ in place of adaptor_align.cpp, barcode_align.cpp, general_align.cpp and reference_align.{h,cpp};
 * src/init.cpp keeps its registration table (src/init.cpp:9-35) unchanged, because the symbols, arities and return shapes
 * are the same.  It depends only on Rinternals.h, Biostrings_interface.h and sarlacc_b200.h.  There is no R in this image:
 * the file is compiled against the stand-in headers of tests/rstub/ and driven through the SEXP layer by
 * tests/rstub/glue_driver.cpp (tests/test_r_glue.py), under AddressSanitizer for the error paths.  See INTEGRATION.md.
 *
 * Error discipline.  Rf_error() longjmp()s to R's top level and runs no C++ destructor on the way.  The reference keeps
 * its std::vector / std::string objects safe with BEGIN_RCPP / END_RCPP (src/adaptor_align.cpp:12,76: a try block whose
 * catch clauses hand the message to R only after the stack has been unwound).  The same here: every entry point runs its
 * body inside guarded(), failures inside the body are C++ exceptions, and Rf_error() is raised from guarded()'s frame
 * after the body's objects are gone.  R allocations that can themselves fail (Rf_allocVector) are made before the body
 * builds any C++ container.
 *
 * Threading: called on R's main thread; SEXPs are only touched here, before and after the library call.  The library
 * reports errors by return code, so no exception or longjmp crosses CUDA resources.
 */
#include <Rinternals.h>
extern "C" {
#include "Biostrings_interface.h"
}
#include "sarlacc_b200.h"

#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

char g_message[1024];      /* outlives the unwinding; R copies it before Rf_error's longjmp leaves this file */

template <class Body>
SEXP guarded(Body&& body) {
    bool failed = false;
    SEXP out = R_NilValue;
    try {
        out = body();
    } catch (const std::exception& e) {
        std::snprintf(g_message, sizeof(g_message), "%s", e.what());
        failed = true;
    } catch (...) {
        std::snprintf(g_message, sizeof(g_message), "C++ exception (unknown reason)");
        failed = true;
    }
    if (failed) Rf_error("%s", g_message);     /* every object of body() has been destroyed */
    return out;
}

[[noreturn]] void stop(const std::string& msg) { throw std::runtime_error(msg); }

void check(int rc) {
    if (rc != 0) stop(sarlacc_last_error());
}

int read_count(SEXP readseq) {      /* no C++ allocation: safe to call before the R outputs exist */
    if (IS_S4_OBJECT(readseq)) {
        XStringSet_holder s = hold_XStringSet(readseq);
        return get_length_from_XStringSet_holder(&s);
    }
    return LENGTH(readseq);
}

struct Reads {   /* views layout: one (pointer, length) pair per XStringSet element, src/adaptor_align.cpp:46-50 */
    std::vector<const uint8_t*> sp, qp;
    std::vector<int32_t> sl, ql;
    sarlacc_reads r;
    Reads(SEXP readseq, SEXP readqual) {
        XStringSet_holder q = hold_XStringSet(readqual);
        const int nq = get_length_from_XStringSet_holder(&q);
        int ns;
        if (IS_S4_OBJECT(readseq)) {            /* DNAStringSet: Biostrings byte codes, src/DNA_input.cpp:64-75 */
            XStringSet_holder s = hold_XStringSet(readseq);
            ns = get_length_from_XStringSet_holder(&s);
            sp.resize(ns); sl.resize(ns);
            for (int i = 0; i < ns; ++i) {
                Chars_holder e = get_elt_from_XStringSet_holder(&s, i);
                sp[i] = reinterpret_cast<const uint8_t*>(e.ptr);
                sl[i] = e.length;
            }
            r.seq_encoding = SARLACC_SEQ_BIOSTRINGS;
        } else {                                 /* character vector, src/DNA_input.cpp:47-51 */
            ns = LENGTH(readseq);
            sp.resize(ns); sl.resize(ns);
            for (int i = 0; i < ns; ++i) {
                SEXP e = STRING_ELT(readseq, i);
                sp[i] = reinterpret_cast<const uint8_t*>(CHAR(e));
                sl[i] = LENGTH(e);
            }
            r.seq_encoding = SARLACC_SEQ_ASCII;
        }
        if (ns != nq) stop("sequence and quality vectors should have the same length");   /* :23-25 */
        qp.resize(nq); ql.resize(nq);
        for (int i = 0; i < nq; ++i) {
            Chars_holder e = get_elt_from_XStringSet_holder(&q, i);
            qp[i] = reinterpret_cast<const uint8_t*>(e.ptr);
            ql[i] = e.length;
        }
        r.n = ns;
        r.seq = sp.data(); r.seq_len = sl.data();
        r.qual = qp.data(); r.qual_len = ql.data();
        r.seq_pool = r.qual_pool = NULL; r.seq_off = r.qual_off = NULL;
    }
};

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
    if (!Rf_isNumeric(x) || LENGTH(x) != 1) stop(std::string(what) + " should be a numeric scalar");
    return Rf_asReal(x);
}

const char* string_scalar(SEXP x, const char* what) { /* src/utils.cpp:26-31 */
    if (!Rf_isString(x) || LENGTH(x) != 1) stop(std::string(what) + " should be a string");
    return CHAR(STRING_ELT(x, 0));
}

}  // namespace

extern "C" {

SEXP adaptor_align(SEXP readseq, SEXP readqual, SEXP encoding, SEXP gapopen, SEXP gapext, SEXP adaptor, SEXP sec_starts, SEXP sec_ends) {
    return guarded([&]() -> SEXP {
        const char* ad = string_scalar(adaptor, "adaptor sequence");
        const double go = numeric_scalar(gapopen, "gap opening penalty"), ge = numeric_scalar(gapext, "gap extension penalty");
        const int nsec = LENGTH(sec_starts);
        if (nsec != LENGTH(sec_ends)) stop("section starts and ends should have the same length");    /* :29-31 */
        const int n = read_count(readseq);
        /* R outputs first: List(scores, starts, ends, List(section starts), List(section widths)), :71-74 */
        SEXP out = PROTECT(Rf_allocVector(VECSXP, 5));
        SEXP score = Rf_allocVector(REALSXP, n); SET_VECTOR_ELT(out, 0, score);
        SEXP start = Rf_allocVector(INTSXP, n);  SET_VECTOR_ELT(out, 1, start);
        SEXP end = Rf_allocVector(INTSXP, n);    SET_VECTOR_ELT(out, 2, end);
        SEXP ss = Rf_allocVector(VECSXP, nsec);  SET_VECTOR_ELT(out, 3, ss);
        SEXP sw = Rf_allocVector(VECSXP, nsec);  SET_VECTOR_ELT(out, 4, sw);
        for (int s = 0; s < nsec; ++s) {
            SET_VECTOR_ELT(ss, s, Rf_allocVector(INTSXP, n));
            SET_VECTOR_ELT(sw, s, Rf_allocVector(INTSXP, n));
        }
        {
            Reads R(readseq, readqual);
            Enc E(encoding);
            std::vector<int32_t> sst((size_t)nsec * n + 1), swd((size_t)nsec * n + 1);   /* [nsec][n], one IntegerVector per section below */
            check(sarlacc_adaptor_align(&R.r, &E.e, go, ge, ad, nsec, INTEGER(sec_starts), INTEGER(sec_ends),
                                        REAL(score), INTEGER(start), INTEGER(end), sst.data(), swd.data()));
            for (int s = 0; s < nsec; ++s) {
                if (n > 0) {
                    std::memcpy(INTEGER(VECTOR_ELT(ss, s)), sst.data() + (size_t)s * n, sizeof(int32_t) * (size_t)n);
                    std::memcpy(INTEGER(VECTOR_ELT(sw, s)), swd.data() + (size_t)s * n, sizeof(int32_t) * (size_t)n);
                }
            }
        }
        UNPROTECT(1);
        return out;
    });
}

static SEXP score_only(SEXP readseq, SEXP readqual, SEXP encoding, SEXP gapopen, SEXP gapext, SEXP ref, const char* what, bool global) {
    return guarded([&]() -> SEXP {
        const char* rf = string_scalar(ref, what);
        const double go = numeric_scalar(gapopen, "gap opening penalty"), ge = numeric_scalar(gapext, "gap extension penalty");
        SEXP score = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)read_count(readseq)));
        {
            Reads R(readseq, readqual);
            Enc E(encoding);
            check(global ? sarlacc_barcode_align(&R.r, &E.e, go, ge, rf, REAL(score))
                         : sarlacc_adaptor_align_score_only(&R.r, &E.e, go, ge, rf, REAL(score)));
        }
        UNPROTECT(1);
        return score;
    });
}

SEXP adaptor_align_score_only(SEXP readseq, SEXP readqual, SEXP encoding, SEXP gapopen, SEXP gapext, SEXP adaptor) {
    return score_only(readseq, readqual, encoding, gapopen, gapext, adaptor, "adaptor sequence", false);
}

SEXP barcode_align(SEXP barcodeseq, SEXP barcodequal, SEXP encoding, SEXP gapopen, SEXP gapext, SEXP reference) {
    return score_only(barcodeseq, barcodequal, encoding, gapopen, gapext, reference, "barcode sequence", true);
}

SEXP general_align(SEXP inputseq, SEXP inputqual, SEXP encoding, SEXP gapopen, SEXP gapext, SEXP reference, SEXP edit_only) {
    return guarded([&]() -> SEXP {
        const char* rf = string_scalar(reference, "reference sequence");
        const double go = numeric_scalar(gapopen, "gap opening penalty"), ge = numeric_scalar(gapext, "gap extension penalty");
        if (!Rf_isLogical(edit_only) || LENGTH(edit_only) != 1) stop("edit-only specification should be a logical scalar");
        const int eo = Rf_asLogical(edit_only);
        const int n = read_count(inputseq);
        SEXP out = PROTECT(Rf_allocVector(VECSXP, 4));
        SEXP score = Rf_allocVector(REALSXP, n); SET_VECTOR_ELT(out, 0, score);
        SEXP edit = Rf_allocVector(INTSXP, n);   SET_VECTOR_ELT(out, 1, edit);
        SEXP rs = Rf_allocVector(STRSXP, eo ? 0 : n); SET_VECTOR_ELT(out, 2, rs);
        SEXP qs = Rf_allocVector(STRSXP, eo ? 0 : n); SET_VECTOR_ELT(out, 3, qs);
        std::vector<char> ra, qa;
        int64_t stride = 0;
        {
            Reads R(inputseq, inputqual);
            Enc E(encoding);
            int maxlen = 0;
            for (int i = 0; i < n; ++i) if (R.sl[i] > maxlen) maxlen = R.sl[i];
            stride = (int64_t)maxlen + (int64_t)std::strlen(rf) + 2;
            ra.assign(eo ? 1 : (size_t)stride * n + 1, 0);
            qa.assign(eo ? 1 : (size_t)stride * n + 1, 0);
            check(sarlacc_general_align(&R.r, &E.e, go, ge, rf, eo, REAL(score), INTEGER(edit), ra.data(), qa.data(), stride));
        }
        if (!eo) {
            /* Rf_mkChar allocates: the gapped strings are the only C++ objects still alive, as in the reference's
             * Rcpp::StringVector fill (src/general_align.cpp:55-57), and guarded() still owns the unwinding */
            for (int i = 0; i < n; ++i) {
                SET_STRING_ELT(rs, i, Rf_mkChar(ra.data() + (size_t)i * stride));
                SET_STRING_ELT(qs, i, Rf_mkChar(qa.data() + (size_t)i * stride));
            }
        }
        UNPROTECT(1);
        return out;
    });
}

}

/* ---- optional extra routines: R-level loops folded into one call each ------------------------------
 * barcode_align_multi(seq, qual, encoding, gapopen, gapext, barcodes): the per-barcode loop of R/barcodeAlign.R:20-35.
 * Returns list(id, best, next): id = 1-based index of the best barcode (strict >, in barcode order, :28-34), its score,
 * the runner-up's score -- what the R loop leaves in `current.id`, `current.score` and `next.best`.
 * adaptor_align_reads(readseq, readqual, tolerance, encoding, gapopen, gapext, adaptor1, adaptor2, starts1, ends1, starts2,
 * ends2): .align_AA_internal (R/adaptorAlign.R:178-199) on whole reads + adaptorAlign's adaptor2 flip (:66-71).  Returns
 * list(reversed, width, adaptor1, adaptor2), the last two shaped like adaptor_align()'s return value. */
extern "C" {

SEXP barcode_align_multi(SEXP barcodeseq, SEXP barcodequal, SEXP encoding, SEXP gapopen, SEXP gapext, SEXP barcodes) {
    return guarded([&]() -> SEXP {
        const double go = numeric_scalar(gapopen, "gap opening penalty"), ge = numeric_scalar(gapext, "gap extension penalty");
        if (!Rf_isString(barcodes) || LENGTH(barcodes) < 1) stop("barcodes should be a non-empty character vector");
        const int n = read_count(barcodeseq), nb = LENGTH(barcodes);
        SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
        SEXP id = Rf_allocVector(INTSXP, n);    SET_VECTOR_ELT(out, 0, id);
        SEXP best = Rf_allocVector(REALSXP, n); SET_VECTOR_ELT(out, 1, best);
        SEXP next = Rf_allocVector(REALSXP, n); SET_VECTOR_ELT(out, 2, next);
        {
            Reads R(barcodeseq, barcodequal);
            Enc E(encoding);
            std::vector<const char*> bc((size_t)nb);
            for (int b = 0; b < nb; ++b) bc[(size_t)b] = CHAR(STRING_ELT(barcodes, b));
            check(sarlacc_barcode_align_multi(&R.r, &E.e, go, ge, bc.data(), nb, INTEGER(id), REAL(best), REAL(next), NULL));
        }
        UNPROTECT(1);
        return out;
    });
}

SEXP adaptor_align_reads(SEXP readseq, SEXP readqual, SEXP tolerance, SEXP encoding, SEXP gapopen, SEXP gapext,
                         SEXP adaptor1, SEXP adaptor2, SEXP starts1, SEXP ends1, SEXP starts2, SEXP ends2) {
    return guarded([&]() -> SEXP {
        const char* a1 = string_scalar(adaptor1, "adaptor sequence");
        const char* a2 = string_scalar(adaptor2, "adaptor sequence");
        const double go = numeric_scalar(gapopen, "gap opening penalty"), ge = numeric_scalar(gapext, "gap extension penalty");
        const int tol = (int)numeric_scalar(tolerance, "tolerance");
        const int ns[2] = {LENGTH(starts1), LENGTH(starts2)};
        if (ns[0] != LENGTH(ends1) || ns[1] != LENGTH(ends2)) stop("section starts and ends should have the same length");
        const int n = read_count(readseq);
        SEXP out = PROTECT(Rf_allocVector(VECSXP, 4));
        SEXP rev = Rf_allocVector(LGLSXP, n);   SET_VECTOR_ELT(out, 0, rev);
        SEXP width = Rf_allocVector(INTSXP, n); SET_VECTOR_ELT(out, 1, width);
        SEXP per[2], score[2], start[2], end[2], ss[2], sw[2];
        for (int k = 0; k < 2; ++k) {
            per[k] = Rf_allocVector(VECSXP, 5);        SET_VECTOR_ELT(out, 2 + k, per[k]);
            score[k] = Rf_allocVector(REALSXP, n);     SET_VECTOR_ELT(per[k], 0, score[k]);
            start[k] = Rf_allocVector(INTSXP, n);      SET_VECTOR_ELT(per[k], 1, start[k]);
            end[k] = Rf_allocVector(INTSXP, n);        SET_VECTOR_ELT(per[k], 2, end[k]);
            ss[k] = Rf_allocVector(VECSXP, ns[k]);     SET_VECTOR_ELT(per[k], 3, ss[k]);
            sw[k] = Rf_allocVector(VECSXP, ns[k]);     SET_VECTOR_ELT(per[k], 4, sw[k]);
            for (int s = 0; s < ns[k]; ++s) {
                SET_VECTOR_ELT(ss[k], s, Rf_allocVector(INTSXP, n));
                SET_VECTOR_ELT(sw[k], s, Rf_allocVector(INTSXP, n));
            }
        }
        {
            Reads R(readseq, readqual);
            Enc E(encoding);
            std::vector<uint8_t> flag((size_t)n + 1);
            std::vector<int32_t> sst[2], swd[2];
            for (int k = 0; k < 2; ++k) {
                sst[k].assign((size_t)ns[k] * n + 1, 0);
                swd[k].assign((size_t)ns[k] * n + 1, 0);
            }
            check(sarlacc_adaptor_align_reads(&R.r, tol, &E.e, go, ge, a1, a2,
                                              ns[0], INTEGER(starts1), INTEGER(ends1), ns[1], INTEGER(starts2), INTEGER(ends2),
                                              INTEGER(width), flag.data(),
                                              REAL(score[0]), INTEGER(start[0]), INTEGER(end[0]), sst[0].data(), swd[0].data(),
                                              REAL(score[1]), INTEGER(start[1]), INTEGER(end[1]), sst[1].data(), swd[1].data()));
            for (int i = 0; i < n; ++i) LOGICAL(rev)[i] = flag[(size_t)i] ? 1 : 0;
            for (int k = 0; k < 2; ++k) {
                for (int s = 0; s < ns[k] && n > 0; ++s) {
                    std::memcpy(INTEGER(VECTOR_ELT(ss[k], s)), sst[k].data() + (size_t)s * n, sizeof(int32_t) * (size_t)n);
                    std::memcpy(INTEGER(VECTOR_ELT(sw[k], s)), swd[k].data() + (size_t)s * n, sizeof(int32_t) * (size_t)n);
                }
            }
        }
        UNPROTECT(1);
        return out;
    });
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
    if (!Rf_isInteger(x) || LENGTH(x) != 1) stop(std::string(what) + " should be an integer scalar");
    return INTEGER(x)[0];
}

struct Lists {
    std::vector<int64_t> off;
    std::vector<int32_t> val;
    explicit Lists(sarlacc_lists* h) {          /* takes the handle */
        if (!h) stop(sarlacc_last_error());
        off.assign((size_t)sarlacc_lists_count(h) + 1, 0);
        val.assign((size_t)sarlacc_lists_values(h) + 1, 0);
        sarlacc_lists_fetch(h, off.data(), val.data());
        sarlacc_lists_free(h);
    }
    /* VECSXP of INTSXP; the caller protects the result */
    SEXP to_sexp() const {
        const R_xlen_t n = (R_xlen_t)off.size() - 1;
        SEXP out = PROTECT(Rf_allocVector(VECSXP, n));
        for (R_xlen_t c = 0; c < n; ++c) {
            SEXP v = Rf_allocVector(INTSXP, (R_xlen_t)(off[c + 1] - off[c]));
            SET_VECTOR_ELT(out, c, v);
            if (off[c + 1] > off[c]) std::memcpy(INTEGER(v), val.data() + off[c], sizeof(int32_t) * (size_t)(off[c + 1] - off[c]));
        }
        UNPROTECT(1);
        return out;
    }
};

}

extern "C" {

SEXP umi_group(SEXP umi1, SEXP thresh1, SEXP umi2, SEXP thresh2, SEXP pregroup) {
    return guarded([&]() -> SEXP {
        const int t1 = integer_scalar(thresh1, "threshold 1");
        const bool two = umi2 != R_NilValue;
        const int t2 = integer_scalar(thresh2, "threshold 2");
        const R_xlen_t ng = Rf_xlength(pregroup);
        SEXP out = PROTECT(Rf_allocVector(VECSXP, ng));
        {
            UmiPool u1(umi1);
            UmiPool u2(two ? umi2 : umi1);
            if (two && u1.n != u2.n) stop("'umi1' and 'umi2' should have the same length");   /* src/umi_group.cpp:25-29 */
            /* one library call per pre-group keeps the reference's nesting (a list per group) without a second index */
            for (R_xlen_t g = 0; g < ng; ++g) {
                SEXP cur = VECTOR_ELT(pregroup, g);
                const int64_t goff[2] = {0, (int64_t)LENGTH(cur)};
                const Lists L(sarlacc_umi_group(u1.pool.data(), u1.off.data(), u1.n, t1, two ? u2.pool.data() : NULL, u2.off.data(), t2,
                                                goff, INTEGER(cur), 1, /*device*/ 0));
                SET_VECTOR_ELT(out, g, L.to_sexp());
            }
        }
        UNPROTECT(1);
        return out;
    });
}
/* (With many small pre-groups, pass them all in ONE call -- group_off / members as CSR -- and re-nest by counting the
 *  clusters per group: the device pass then covers every group at once.  sarlacc_b200/native.py: umi_group does that.) */

SEXP cluster_umis_test(SEXP links) {
    return guarded([&]() -> SEXP {
        const R_xlen_t n = Rf_xlength(links);
        std::vector<int64_t> loff(1, 0);
        std::vector<int32_t> lval;
        for (R_xlen_t i = 0; i < n; ++i) {
            SEXP cur = VECTOR_ELT(links, i);
            lval.insert(lval.end(), INTEGER(cur), INTEGER(cur) + LENGTH(cur));
            loff.push_back((int64_t)lval.size());
        }
        if (lval.empty()) lval.push_back(0);
        const Lists L(sarlacc_cluster_umis(loff.data(), lval.data(), (int64_t)n));
        return L.to_sexp();
    });
}

}
