/* TEST INFRASTRUCTURE ONLY -- not part of the product.
 *
 * Common C interface of the two CPU oracles:
 *   ref_*  : oracle/_ref/libsarlacc_ref.so  -- the reference's own reference_align.cpp /
 *            quality_encoding.cpp compiled verbatim from /root/reference (oracle/ref_driver.cpp
 *            restates only the per-read entry loops of src/adaptor_align.cpp:45-69,96-106,
 *            src/barcode_align.cpp:29-40, src/general_align.cpp:34-58).
 *   orc_*  : oracle/libsarlacc_oracle.so    -- plain-C restatement (oracle/sarlacc_oracle.c).
 *
 * Inputs are CSR: a byte pool plus n+1 offsets, separately for sequences and qualities (so that
 * the "sequence and quality strings should have the same length" error path can be exercised).
 * Every function returns 0 on success, 1 on error with the reference's message copied into `err`.
 * With nthreads > 1 reads are split into contiguous chunks like .parallelize (R/adaptorAlign.R:126-134);
 * the error reported is the one of the lowest read index, as a serial run would report.
 */
#ifndef SARLACC_ORACLE_ABI_H
#define SARLACC_ORACLE_ABI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_DECL(P) \
int P##_adaptor_align(int64_t n, const char* seq, const int64_t* seq_off, const char* qual, const int64_t* qual_off, \
        int enc_n, const char* const* enc_names, const double* enc_err, double go, double ge, \
        const char* adaptor, int nsec, const int32_t* sec_starts, const int32_t* sec_ends, \
        double* score, int32_t* start, int32_t* end, int32_t* sec_start, int32_t* sec_width, \
        int nthreads, char* err, int errlen); \
int P##_align_score_only(int64_t n, const char* seq, const int64_t* seq_off, const char* qual, const int64_t* qual_off, \
        int enc_n, const char* const* enc_names, const double* enc_err, double go, double ge, \
        const char* reference, int local, double* score, int nthreads, char* err, int errlen); \
int P##_general_align(int64_t n, const char* seq, const int64_t* seq_off, const char* qual, const int64_t* qual_off, \
        int enc_n, const char* const* enc_names, const double* enc_err, double go, double ge, \
        const char* reference, int edit_only, double* score, int32_t* edit, \
        char* ref_aln, char* query_aln, int64_t aln_stride, int nthreads, char* err, int errlen); \
int P##_cost_tables(int enc_n, const char* const* enc_names, const double* enc_err, \
        double* match /*[4][enc_n]*/, double* mismatch /*[4][enc_n]*/, char* offset, char* err, int errlen);

ORACLE_DECL(ref)
ORACLE_DECL(orc)

#ifdef __cplusplus
}
#endif

#endif
