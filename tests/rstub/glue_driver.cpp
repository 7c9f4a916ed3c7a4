/* TEST INFRASTRUCTURE ONLY -- drives sarlacc_b200/csrc/r_glue.cpp through the SEXP layer of tests/rstub/ (no R needed).
 *
 *   glue_driver errors            every error the glue can raise without a device, each through Rf_error's longjmp; built
 *                                 with AddressSanitizer, so an error raised while C++ objects are alive shows up as a leak
 *   glue_driver gpu_errors        the reference's per-read errors, which the library finds on the device path
 *   glue_driver parity <reads>    a fixed battery of calls on the reads of a file ("SEQ QUAL" per line); results as JSON
 *                                 lines for tests/test_r_glue.py to compare with the library's own Python face
 */
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "rstub.h"

extern "C" {
SEXP adaptor_align(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP adaptor_align_score_only(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP barcode_align(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP general_align(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP umi_group(SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP cluster_umis_test(SEXP);
SEXP barcode_align_multi(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP adaptor_align_reads(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
}

namespace {

std::vector<std::string> g_names;
std::vector<double> g_err;

SEXP phred_encoding(bool named = true) {
    if (g_names.empty()) {
        for (int q = 0; q < 94; ++q) {
            g_names.push_back(std::string(1, (char)(33 + q)));
            g_err.push_back(std::pow(10.0, -q / 10.0));
        }
    }
    std::vector<const char*> nm;
    for (auto& s : g_names) nm.push_back(s.c_str());
    return rstub_named_reals(g_err.data(), named ? nm.data() : nullptr, 94);
}

std::vector<const char*> ptrs(const std::vector<std::string>& v) {
    std::vector<const char*> p;
    for (auto& s : v) p.push_back(s.c_str());
    if (p.empty()) p.push_back("");
    return p;
}

int failures = 0;

/* Runs `call` under the toy top level; returns the R error message ("" if none). */
template <class F>
std::string run(F&& call, SEXP* result = nullptr) {
    std::string msg;
    if (setjmp(rstub_top_level) == 0) {
        SEXP r = call();
        if (result) *result = r;
        if (rstub_protect_depth() != 0) {
            std::printf("{\"problem\": \"unbalanced PROTECT: %d\"}\n", rstub_protect_depth());
            ++failures;
        }
    } else {
        msg = rstub_last_error();
        rstub_reset_protect();
    }
    return msg;
}

void expect_error(const char* name, const std::string& got, const char* want) {
    const bool ok = got == want;
    std::printf("{\"case\": \"%s\", \"error\": \"%s\", \"ok\": %s}\n", name, got.c_str(), ok ? "true" : "false");
    if (!ok) ++failures;
}

void print_doubles(const char* key, SEXP x) {
    std::printf("\"%s\": [", key);
    for (int i = 0; i < LENGTH(x); ++i) std::printf("%s\"%a\"", i ? ", " : "", REAL(x)[i]);
    std::printf("]");
}

void print_ints(const char* key, SEXP x) {
    std::printf("\"%s\": [", key);
    for (int i = 0; i < LENGTH(x); ++i) std::printf("%s%d", i ? ", " : "", INTEGER(x)[i]);
    std::printf("]");
}

void print_int_lists(const char* key, SEXP x) {
    std::printf("\"%s\": [", key);
    for (int k = 0; k < LENGTH(x); ++k) {
        SEXP v = VECTOR_ELT(x, k);
        std::printf("%s[", k ? ", " : "");
        for (int i = 0; i < LENGTH(v); ++i) std::printf("%s%d", i ? ", " : "", INTEGER(v)[i]);
        std::printf("]");
    }
    std::printf("]");
}

void print_strings(const char* key, SEXP x) {
    std::printf("\"%s\": [", key);
    for (int i = 0; i < LENGTH(x); ++i) std::printf("%s\"%s\"", i ? ", " : "", CHAR(STRING_ELT(x, i)));
    std::printf("]");
}

int errors_mode(bool have_gpu) {
    const std::vector<std::string> seqs = {"ACGTACGTAAGG", "ACGTTTGGA"}, quals = {"555555555555", "555555555"};
    auto sp = ptrs(seqs), qp = ptrs(quals);
    const int one = 1, two[2] = {1, 2};
    auto reads = [&] { return rstub_string_vector(sp.data(), 2); };
    auto rqual = [&] { return rstub_xstringset(qp.data(), 2, 0); };
    if (!have_gpu) {
        expect_error("adaptor not a string", run([&] { return adaptor_align(reads(), rqual(), phred_encoding(), rstub_real(5), rstub_real(1), rstub_real(3), rstub_integers(&one, 0), rstub_integers(&one, 0)); }),
                     "adaptor sequence should be a string");
        expect_error("gap opening not scalar", run([&] { const double v[2] = {5, 6}; return adaptor_align(reads(), rqual(), phred_encoding(), rstub_named_reals(v, nullptr, 2), rstub_real(1), rstub_string("ACGT"), rstub_integers(&one, 0), rstub_integers(&one, 0)); }),
                     "gap opening penalty should be a numeric scalar");
        expect_error("gap extension not numeric", run([&] { return adaptor_align_score_only(reads(), rqual(), phred_encoding(), rstub_real(5), rstub_string("x"), rstub_string("ACGT")); }),
                     "gap extension penalty should be a numeric scalar");
        expect_error("section lengths differ", run([&] { return adaptor_align(reads(), rqual(), phred_encoding(), rstub_real(5), rstub_real(1), rstub_string("ACGT"), rstub_integers(two, 2), rstub_integers(&one, 1)); }),
                     "section starts and ends should have the same length");
        expect_error("vector lengths differ", run([&] { return adaptor_align(reads(), rstub_xstringset(qp.data(), 1, 0), phred_encoding(), rstub_real(5), rstub_real(1), rstub_string("ACGT"), rstub_integers(&one, 0), rstub_integers(&one, 0)); }),
                     "sequence and quality vectors should have the same length");
        expect_error("vector lengths differ (score only)", run([&] { return barcode_align(reads(), rstub_xstringset(qp.data(), 1, 0), phred_encoding(), rstub_real(5), rstub_real(1), rstub_string("ACGT")); }),
                     "sequence and quality vectors should have the same length");
        expect_error("barcode not a string", run([&] { return barcode_align(reads(), rqual(), phred_encoding(), rstub_real(5), rstub_real(1), rstub_string_vector(sp.data(), 2)); }),
                     "barcode sequence should be a string");
        expect_error("edit_only not logical", run([&] { return general_align(reads(), rqual(), phred_encoding(), rstub_real(5), rstub_real(1), rstub_string("ACGT"), rstub_real(1)); }),
                     "edit-only specification should be a logical scalar");
        expect_error("unnamed encoding", run([&] { return adaptor_align(reads(), rqual(), phred_encoding(false), rstub_real(5), rstub_real(1), rstub_string("ACGT"), rstub_integers(&one, 0), rstub_integers(&one, 0)); }),
                     "encoding vector must be non-empty and named");
        {
            const double e[3] = {0.1, 0.01, 0.001};
            const char* n1[3] = {"!", "\"\"", "#"};
            expect_error("encoding names too long", run([&] { return general_align(reads(), rqual(), rstub_named_reals(e, n1, 3), rstub_real(5), rstub_real(1), rstub_string("ACGT"), rstub_logical(1)); }),
                         "names of encoding vector must be one character in length");
            const char* n2[3] = {"!", "#", "$"};
            expect_error("encoding names not consecutive", run([&] { return adaptor_align_score_only(reads(), rqual(), rstub_named_reals(e, n2, 3), rstub_real(5), rstub_real(1), rstub_string("ACGT")); }),
                         "names of encoding vector should increase consecutively");
            const double e2[3] = {0.1, 0.2, 0.001};
            const char* n3[3] = {"!", "\"", "#"};
            expect_error("encoding errors increase", run([&] { return barcode_align(reads(), rqual(), rstub_named_reals(e2, n3, 3), rstub_real(5), rstub_real(1), rstub_string("ACGT")); }),
                         "error probabilities should decrease");
        }
        {
            const int t = 1;
            const int g0[2] = {1, 2};
            SEXP err_msg_holder = nullptr;
            (void)err_msg_holder;
            expect_error("umi threshold not integer", run([&] { SEXP pg = rstub_list(1); SET_VECTOR_ELT(pg, 0, rstub_integers(g0, 2)); return umi_group(reads(), rstub_real(1), R_NilValue, rstub_integers(&t, 1), pg); }),
                         "threshold 1 should be an integer scalar");
            expect_error("umi vectors differ", run([&] { SEXP pg = rstub_list(1); SET_VECTOR_ELT(pg, 0, rstub_integers(g0, 2)); return umi_group(reads(), rstub_integers(&t, 1), rstub_string_vector(sp.data(), 1), rstub_integers(&t, 1), pg); }),
                         "'umi1' and 'umi2' should have the same length");
        }
        /* the optional fused routines */
        expect_error("barcodes not character", run([&] { return barcode_align_multi(reads(), rqual(), phred_encoding(), rstub_real(5), rstub_real(1), rstub_real(2)); }),
                     "barcodes should be a non-empty character vector");
        expect_error("vector lengths differ (multi)", run([&] { return barcode_align_multi(reads(), rstub_xstringset(qp.data(), 1, 0), phred_encoding(), rstub_real(5), rstub_real(1), rstub_string_vector(sp.data(), 2)); }),
                     "sequence and quality vectors should have the same length");
        expect_error("tolerance not numeric", run([&] { return adaptor_align_reads(reads(), rqual(), rstub_string("x"), phred_encoding(), rstub_real(5), rstub_real(1), rstub_string("ACGT"), rstub_string("TTGA"), rstub_integers(&one, 0), rstub_integers(&one, 0), rstub_integers(&one, 0), rstub_integers(&one, 0)); }),
                     "tolerance should be a numeric scalar");
        expect_error("section lengths differ (reads)", run([&] { return adaptor_align_reads(reads(), rqual(), rstub_real(250), phred_encoding(), rstub_real(5), rstub_real(1), rstub_string("ACGT"), rstub_string("TTGA"), rstub_integers(&one, 0), rstub_integers(&one, 0), rstub_integers(two, 2), rstub_integers(&one, 1)); }),
                     "section starts and ends should have the same length");
        expect_error("second adaptor not a string", run([&] { return adaptor_align_reads(reads(), rqual(), rstub_real(250), phred_encoding(), rstub_real(5), rstub_real(1), rstub_string("ACGT"), rstub_real(1), rstub_integers(&one, 0), rstub_integers(&one, 0), rstub_integers(&one, 0), rstub_integers(&one, 0)); }),
                     "adaptor sequence should be a string");
        /* a valid call without a device: the library's own refusal travels the same way, with every container alive */
        const std::string nodev = run([&] { return adaptor_align(reads(), rqual(), phred_encoding(), rstub_real(5), rstub_real(1), rstub_string("ACGT"), rstub_integers(&one, 0), rstub_integers(&one, 0)); });
        const bool refused = nodev.find("requires a CUDA device") != std::string::npos;
        std::printf("{\"case\": \"no device\", \"error\": \"%s\", \"ok\": %s}\n", nodev.c_str(), (refused || nodev.empty()) ? "true" : "false");
        if (!refused && !nodev.empty()) ++failures;
        /* cluster_umis_test runs on the host: a result, not an error */
        {
            const int l0[2] = {1, 2}, l1[2] = {1, 2}, l2[1] = {3};
            SEXP res = nullptr;
            const std::string m = run([&] { SEXP links = rstub_list(3); SET_VECTOR_ELT(links, 0, rstub_integers(l0, 2)); SET_VECTOR_ELT(links, 1, rstub_integers(l1, 2)); SET_VECTOR_ELT(links, 2, rstub_integers(l2, 1)); return cluster_umis_test(links); }, &res);
            std::printf("{\"case\": \"cluster_umis_test\", \"error\": \"%s\", \"clusters\": %d, \"ok\": %s}\n", m.c_str(), res ? LENGTH(res) : -1, (m.empty() && res && LENGTH(res) == 2) ? "true" : "false");
            if (!m.empty() || !res || LENGTH(res) != 2) ++failures;
        }
    } else {
        const std::vector<std::string> bq = {"555555555555", "55555555"};      /* second quality string one short */
        auto bqp = ptrs(bq);
        expect_error("string lengths differ", run([&] { return adaptor_align(reads(), rstub_xstringset(bqp.data(), 2, 0), phred_encoding(), rstub_real(5), rstub_real(1), rstub_string("ACGT"), rstub_integers(&one, 0), rstub_integers(&one, 0)); }),
                     "sequence and quality strings should have the same length");
        const std::vector<std::string> lowq = {"555555555555", "55555 555"};   /* ' ' < '!' */
        auto lqp = ptrs(lowq);
        expect_error("quality below offset", run([&] { return adaptor_align_score_only(reads(), rstub_xstringset(lqp.data(), 2, 0), phred_encoding(), rstub_real(5), rstub_real(1), rstub_string("ACGT")); }),
                     "quality cannot be lower than smallest encoded value");
        expect_error("bad reference base", run([&] { return barcode_align(reads(), rqual(), phred_encoding(), rstub_real(5), rstub_real(1), rstub_string("ACXT")); }),
                     "unrecognized base in reference sequence");
        expect_error("bad reference base (general)", run([&] { return general_align(reads(), rqual(), phred_encoding(), rstub_real(5), rstub_real(1), rstub_string("AC?T"), rstub_logical(0)); }),
                     "unrecognized base in reference sequence");
    }
    rstub_free_all();
    std::printf("{\"failures\": %d}\n", failures);
    return failures ? 1 : 0;
}

int parity_mode(const char* path) {
    std::vector<std::string> seqs, quals;
    std::ifstream in(path);
    std::string s, q;
    while (in >> s >> q) {
        seqs.push_back(s == "-" ? "" : s);
        quals.push_back(q == "-" ? "" : q);
    }
    auto sp = ptrs(seqs), qp = ptrs(quals);
    const int n = (int)seqs.size();
    const char* A1 = "ACGCAGATCGATCGATNNNNNNNNNNNNCGCGCGAGCTGACTNNNNGCACGACTCTGGTTTTTTTTTTTT";
    const int st[2] = {16, 42}, en[2] = {28, 46};
    for (int s4 = 0; s4 < 2; ++s4) {      /* character vector, then DNAStringSet byte codes */
        SEXP res = nullptr;
        const std::string m = run([&] {
            SEXP rs = s4 ? rstub_xstringset(sp.data(), n, 1) : rstub_string_vector(sp.data(), n);
            return adaptor_align(rs, rstub_xstringset(qp.data(), n, 0), phred_encoding(), rstub_real(5), rstub_real(1), rstub_string(A1), rstub_integers(st, 2), rstub_integers(en, 2));
        }, &res);
        if (!m.empty()) { std::printf("{\"call\": \"adaptor_align\", \"error\": \"%s\"}\n", m.c_str()); return 1; }
        std::printf("{\"call\": \"adaptor_align\", \"s4\": %d, ", s4);
        print_doubles("score", VECTOR_ELT(res, 0)); std::printf(", ");
        print_ints("start", VECTOR_ELT(res, 1)); std::printf(", ");
        print_ints("end", VECTOR_ELT(res, 2)); std::printf(", ");
        print_int_lists("sec_start", VECTOR_ELT(res, 3)); std::printf(", ");
        print_int_lists("sec_width", VECTOR_ELT(res, 4)); std::printf("}\n");
    }
    {
        SEXP res = nullptr;
        run([&] { return adaptor_align_score_only(rstub_xstringset(sp.data(), n, 1), rstub_xstringset(qp.data(), n, 0), phred_encoding(), rstub_real(4), rstub_real(2), rstub_string("AAGGCCTTTTCCGACTCATGAA")); }, &res);
        std::printf("{\"call\": \"adaptor_align_score_only\", "); print_doubles("score", res); std::printf("}\n");
        run([&] { return barcode_align(rstub_string_vector(sp.data(), n), rstub_xstringset(qp.data(), n, 0), phred_encoding(), rstub_real(5), rstub_real(1), rstub_string("AAGGCCTTTTCCGACTCATGAACC")); }, &res);
        std::printf("{\"call\": \"barcode_align\", "); print_doubles("score", res); std::printf("}\n");
        const int m = n < 40 ? n : 40;
        run([&] { return general_align(rstub_string_vector(sp.data(), m), rstub_xstringset(qp.data(), m, 0), phred_encoding(), rstub_real(4), rstub_real(1), rstub_string("AAGGAATTAAGGCCTTACGT"), rstub_logical(0)); }, &res);
        std::printf("{\"call\": \"general_align\", "); print_doubles("score", VECTOR_ELT(res, 0)); std::printf(", ");
        print_ints("edit", VECTOR_ELT(res, 1)); std::printf(", "); print_strings("ref", VECTOR_ELT(res, 2)); std::printf(", ");
        print_strings("query", VECTOR_ELT(res, 3)); std::printf("}\n");
        /* umi_group on 12-base prefixes of the reads, two pre-groups */
        std::vector<std::string> umis;
        for (int i = 0; i < n; ++i) umis.push_back(seqs[i].size() >= 12 ? seqs[i].substr(0, 12) : std::string("ACGTACGTACGT"));
        auto up = ptrs(umis);
        std::vector<int> g0, g1;
        for (int i = 0; i < n; ++i) (i % 2 ? g1 : g0).push_back(i + 1);
        const int t = 1;
        run([&] { SEXP pg = rstub_list(2); SET_VECTOR_ELT(pg, 0, rstub_integers(g0.data(), (int)g0.size())); SET_VECTOR_ELT(pg, 1, rstub_integers(g1.data(), (int)g1.size()));
                  return umi_group(rstub_string_vector(up.data(), n), rstub_integers(&t, 1), R_NilValue, rstub_integers(&t, 1), pg); }, &res);
        std::printf("{\"call\": \"umi_group\", \"groups\": [");
        for (int g = 0; g < LENGTH(res); ++g) { std::printf("%s{", g ? ", " : ""); print_int_lists("clusters", VECTOR_ELT(res, g)); std::printf("}"); }
        std::printf("]}\n");
        /* the optional fused routines: three barcodes in one call; both adaptors on both ends of the whole reads */
        const char* bcs[3] = {"AAGGCCTTTTCCGACTCATGAACC", "ACGTACGTACGTACGTACGTACGT", "TTGACCAGTTGACCAGTTGACCAG"};
        run([&] { return barcode_align_multi(rstub_string_vector(sp.data(), n), rstub_xstringset(qp.data(), n, 0), phred_encoding(), rstub_real(5), rstub_real(1), rstub_string_vector(bcs, 3)); }, &res);
        std::printf("{\"call\": \"barcode_align_multi\", "); print_ints("id", VECTOR_ELT(res, 0)); std::printf(", ");
        print_doubles("best", VECTOR_ELT(res, 1)); std::printf(", "); print_doubles("next", VECTOR_ELT(res, 2)); std::printf("}\n");
        const std::string m2 = run([&] { return adaptor_align_reads(rstub_xstringset(sp.data(), n, 1), rstub_xstringset(qp.data(), n, 0), rstub_real(100), phred_encoding(), rstub_real(5), rstub_real(1),
                                              rstub_string(A1), rstub_string("AAGGCCTTTTCCGACTCATGAA"), rstub_integers(st, 2), rstub_integers(en, 2), rstub_integers(st, 0), rstub_integers(en, 0)); }, &res);
        if (!m2.empty()) { std::printf("{\"call\": \"adaptor_align_reads\", \"error\": \"%s\"}\n", m2.c_str()); return 1; }
        std::printf("{\"call\": \"adaptor_align_reads\", "); print_ints("reversed", VECTOR_ELT(res, 0)); std::printf(", "); print_ints("width", VECTOR_ELT(res, 1));
        for (int k = 0; k < 2; ++k) {
            SEXP a = VECTOR_ELT(res, 2 + k);
            std::printf(", \"adaptor%d\": {", k + 1);
            print_doubles("score", VECTOR_ELT(a, 0)); std::printf(", "); print_ints("start", VECTOR_ELT(a, 1)); std::printf(", "); print_ints("end", VECTOR_ELT(a, 2)); std::printf(", ");
            print_int_lists("sec_start", VECTOR_ELT(a, 3)); std::printf(", "); print_int_lists("sec_width", VECTOR_ELT(a, 4)); std::printf("}");
        }
        std::printf("}\n");
    }
    rstub_free_all();
    return 0;
}

/* What the harness is for: an entry point that raises the R error while a container is alive (the pattern the first
 * version of the glue had).  Under AddressSanitizer this run must end with a leak report. */
extern "C" SEXP leaky_entry(SEXP x) {
    std::vector<double> scratch(1000, 1.0);
    if (LENGTH(x) != 3) Rf_error("leaky_entry: %d elements at %p", LENGTH(x), (void*)scratch.data());
    return x;
}

int canary_mode() {
    const int v = 1;
    const std::string m = run([&] { return leaky_entry(rstub_integers(&v, 1)); });
    std::printf("{\"case\": \"canary\", \"error\": \"%s\"}\n", m.c_str());
    std::fflush(stdout);        /* the leak report ends the process without flushing */
    rstub_free_all();
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    const std::string mode = argc > 1 ? argv[1] : "errors";
    if (mode == "canary") return canary_mode();
    if (mode == "errors") return errors_mode(false);
    if (mode == "gpu_errors") return errors_mode(true);
    if (mode == "parity" && argc > 2) return parity_mode(argv[2]);
    std::fprintf(stderr, "usage: glue_driver errors | gpu_errors | parity <reads file>\n");
    return 2;
}
