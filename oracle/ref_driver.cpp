/* TEST INFRASTRUCTURE ONLY -- not part of the product.
 *
 * Drives the reference's OWN DP class (compiled verbatim from /root/reference/src/reference_align.cpp
 * and quality_encoding.cpp against oracle/rcpp_shim/Rcpp.h) through restatements of the reference's
 * per-read entry loops.  Nothing here is shipped or called by the product path; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load the resulting
 * oracle/_ref/libsarlacc_ref.so.
 *
 * Loops restated (reference file:line):
 *   ref_adaptor_align      src/adaptor_align.cpp:11-77   (align -> fill_map -> qmap(0,rlen) guard -> sections)
 *   ref_align_score_only   src/adaptor_align.cpp:79-110 (local=1), src/barcode_align.cpp:10-44 (local=0)
 *   ref_general_align      src/general_align.cpp:10-62
 * Threading (nthreads>1) mirrors .parallelize's contiguous chunks (R/adaptorAlign.R:126-134): one
 * reference_align object per chunk, exactly what one BiocParallel worker would own.
 */
#include "Rcpp.h"
#include <deque>
#include <vector>
#define private public /* cost tables of reference_align are private; the oracle exposes them for table parity */
#include "reference_align.h"
#undef private

#include "oracle_abi.h"

#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace {

Rcpp::NumericVector make_encoding(int enc_n, const char* const* enc_names, const double* enc_err) {
    std::vector<std::string> nm;
    for (int i = 0; i < enc_n; ++i) nm.push_back(enc_names ? std::string(enc_names[i]) : std::string());
    if (!enc_names) nm.clear();
    return Rcpp::NumericVector(enc_err, static_cast<size_t>(enc_n), nm);
}

void put_err(char* err, int errlen, const std::string& msg) {
    if (err && errlen > 0) {
        std::strncpy(err, msg.c_str(), errlen - 1);
        err[errlen - 1] = '\0';
    }
}

/* Runs body(RA-capable lambda) over contiguous chunks; serial-equivalent error reporting. */
template <class F>
int run_chunks(int64_t n, int nthreads, char* err, int errlen, F make_and_run) {
    if (nthreads < 1) nthreads = 1;
    if (static_cast<int64_t>(nthreads) > n) nthreads = n > 0 ? static_cast<int>(n) : 1;
    std::vector<int64_t> err_at(nthreads, -1);
    std::vector<std::string> err_msg(nthreads);
    auto worker = [&](int t) {
        const int64_t lo = n * t / nthreads, hi = n * (t + 1) / nthreads;
        int64_t at = lo;
        try {
            make_and_run(lo, hi, at);
        } catch (std::exception& e) {
            err_at[t] = at;
            err_msg[t] = e.what();
        }
    };
    if (nthreads == 1) {
        worker(0);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; ++t) pool.emplace_back(worker, t);
        for (auto& th : pool) th.join();
    }
    for (int t = 0; t < nthreads; ++t) {
        if (err_at[t] >= 0) {
            put_err(err, errlen, err_msg[t]);
            return 1;
        }
    }
    return 0;
}

}

extern "C" {

int ref_adaptor_align(int64_t n, const char* seq, const int64_t* seq_off, const char* qual, const int64_t* qual_off,
        int enc_n, const char* const* enc_names, const double* enc_err, double go, double ge,
        const char* adaptor, int nsec, const int32_t* sec_starts, const int32_t* sec_ends,
        double* score, int32_t* start, int32_t* end, int32_t* sec_start, int32_t* sec_width,
        int nthreads, char* err, int errlen)
{
    const std::string adaptor_seq(adaptor);
    std::fill(start, start + n, 0);
    std::fill(end, end + n, 0);
    return run_chunks(n, nthreads, err, errlen, [&](int64_t lo, int64_t hi, int64_t& at) {
        /* The constructor validates the encoding (and may throw) before any read is touched. */
        at = 0;
        reference_align RA(adaptor_seq.size(), adaptor_seq.c_str(), make_encoding(enc_n, enc_names, enc_err), go, ge);
        reference_align::querymap qmap;
        for (int64_t i = lo; i < hi; ++i) {
            at = i;
            const char* sstr = seq + seq_off[i];
            const size_t slen = seq_off[i + 1] - seq_off[i];
            if (slen != static_cast<size_t>(qual_off[i + 1] - qual_off[i])) {
                throw std::runtime_error("sequence and quality strings should have the same length");
            }
            score[i] = RA.align(slen, sstr, qual + qual_off[i]);
            RA.fill_map(qmap);

            auto aln_pos = qmap(0, adaptor_seq.size());
            if (aln_pos.first < aln_pos.second) {
                start[i] = aln_pos.first + 1;
                end[i] = aln_pos.second;
            }
            for (int sec = 0; sec < nsec; ++sec) {
                auto current = qmap(sec_starts[sec], sec_ends[sec], true);
                sec_start[static_cast<int64_t>(sec) * n + i] = current.first + 1;
                sec_width[static_cast<int64_t>(sec) * n + i] = current.second - current.first;
            }
        }
    });
}

int ref_align_score_only(int64_t n, const char* seq, const int64_t* seq_off, const char* qual, const int64_t* qual_off,
        int enc_n, const char* const* enc_names, const double* enc_err, double go, double ge,
        const char* reference, int local, double* score, int nthreads, char* err, int errlen)
{
    const std::string ref_seq(reference);
    return run_chunks(n, nthreads, err, errlen, [&](int64_t lo, int64_t hi, int64_t& at) {
        at = 0;
        reference_align RA(ref_seq.size(), ref_seq.c_str(), make_encoding(enc_n, enc_names, enc_err), go, ge);
        for (int64_t i = lo; i < hi; ++i) {
            at = i;
            const size_t slen = seq_off[i + 1] - seq_off[i];
            if (slen != static_cast<size_t>(qual_off[i + 1] - qual_off[i])) {
                throw std::runtime_error("sequence and quality strings should have the same length");
            }
            score[i] = RA.align(slen, seq + seq_off[i], qual + qual_off[i], local != 0);
        }
    });
}

int ref_general_align(int64_t n, const char* seq, const int64_t* seq_off, const char* qual, const int64_t* qual_off,
        int enc_n, const char* const* enc_names, const double* enc_err, double go, double ge,
        const char* reference, int edit_only, double* score, int32_t* edit,
        char* ref_aln, char* query_aln, int64_t aln_stride, int nthreads, char* err, int errlen)
{
    const std::string ref_seq(reference);
    std::fill(edit, edit + n, 0);
    return run_chunks(n, nthreads, err, errlen, [&](int64_t lo, int64_t hi, int64_t& at) {
        at = 0;
        reference_align RA(ref_seq.size(), ref_seq.c_str(), make_encoding(enc_n, enc_names, enc_err), go, ge);
        std::vector<char> tmpref, tmpquery;
        for (int64_t i = lo; i < hi; ++i) {
            at = i;
            const char* sstr = seq + seq_off[i];
            const size_t slen = seq_off[i + 1] - seq_off[i];
            if (slen != static_cast<size_t>(qual_off[i + 1] - qual_off[i])) {
                throw std::runtime_error("sequence and quality strings should have the same length");
            }
            score[i] = RA.align(slen, sstr, qual + qual_off[i], false);
            RA.fill_strings(tmpref, tmpquery, sstr);
            int32_t ed = 0;
            for (size_t j = 0; j < tmpref.size(); ++j) {
                if (tmpref[j] != tmpquery[j]) ++ed;
            }
            edit[i] = ed;
            if (!edit_only) {
                if (static_cast<int64_t>(tmpref.size()) > aln_stride) throw std::runtime_error("oracle: aln_stride too small");
                std::memcpy(ref_aln + i * aln_stride, tmpref.data(), tmpref.size());
                std::memcpy(query_aln + i * aln_stride, tmpquery.data(), tmpquery.size());
            }
        }
    });
}

int ref_cost_tables(int enc_n, const char* const* enc_names, const double* enc_err,
        double* match, double* mismatch, char* offset, char* err, int errlen)
{
    try {
        reference_align RA(0, "", make_encoding(enc_n, enc_names, enc_err), 5, 1);
        for (int m = 0; m < 4; ++m) {
            for (size_t j = 0; j < RA.available; ++j) {
                match[m * enc_n + j] = RA.precomputed_match[m][j];
                mismatch[m * enc_n + j] = RA.precomputed_mismatch[m][j];
            }
        }
        *offset = RA.offset;
    } catch (std::exception& e) {
        put_err(err, errlen, e.what());
        return 1;
    }
    return 0;
}

}
