/* TEST INFRASTRUCTURE ONLY -- not part of the product.
 *
 * Plain-C CPU restatement of the reference's adaptor-alignment hot path (FP64, same operation
 * order and comparison operators), used as the checker in tests/, __graft_entry__.smoke() and as
 * bench.py's cpu_baseline "port" leg.  The product (sarlacc_b200/) never links or calls this.
 *
 * Parity is PINNED: tests/test_oracle.py checks this file bit-for-bit against
 *   (1) oracle/_ref/libsarlacc_ref.so, the reference's own reference_align.cpp compiled verbatim,
 *   (2) the literal known answers of the reference's tests (tests/testthat/test-adaptor-align.R:48-56,
 *       120-127,186-206) and SURVEY.md 8(a)'s golden vectors,
 *   (3) tests/golden/ fixtures generated from (1) by tests/golden/make_golden.py.
 *
 * Each function cites the reference file:line it follows (paths relative to /root/reference).
 * Build: gcc -std=c11 -O2 -ffp-contract=off (no FMA contraction: the reference is built by R's
 * default flags for plain x86-64, which has no FMA instructions to contract into).
 */
#include "oracle_abi.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_LN2
#define M_LN2 0.693147180559945309417232121458176568
#endif

typedef struct {
    /* src/reference_align.h:26-38 */
    size_t rlen;
    const char* rseq;
    double gap_open, gap_ext;
    double* match[4];
    double* mismatch[4];
    char offset;
    size_t available;
    /* src/reference_align.h:40-45; flat arrays instead of std::deque */
    size_t nrows, cap_rows;
    double *scores, *left_jump_scores;
    size_t* left_jump_points;
    int* directions;
    const char* error; /* set instead of throwing */
} aligner;

/* src/quality_encoding.cpp:5-32 */
static const char* check_encoding(int enc_n, const char* const* enc_names, const double* enc_err, char* offset) {
    if (enc_names == NULL || enc_n == 0) {
        return "encoding vector must be non-empty and named";
    }
    char last = 0;
    *offset = 0;
    for (int i = 0; i < enc_n; ++i) {
        if (strlen(enc_names[i]) != 1) {
            return "names of encoding vector must be one character in length";
        }
        const char curval = enc_names[i][0];
        if (i > 0) {
            if (curval != last + 1) { /* signed char arithmetic promoted to int, as in the reference (:20) */
                return "names of encoding vector should increase consecutively";
            } else if (enc_err[i] > enc_err[i - 1]) {
                return "error probabilities should decrease";
            }
        } else {
            *offset = curval;
        }
        last = curval;
    }
    return NULL;
}

/* src/reference_align.cpp:21-52 */
static void fill_tables(int enc_n, const double* errors, double* match, double* mismatch) {
    for (size_t i = 0; i < 4; ++i) {
        double gamma_xy = 1.0 / (i + 1.0);
        double gamma_xy_1m = 1 - gamma_xy;
        const double n = 4;
        for (int j = 0; j < enc_n; ++j) {
            const double epsilon = errors[j];
            match[i * enc_n + j] = log(gamma_xy * (1 - epsilon) * n + gamma_xy_1m * epsilon * (n / (n - 1))) / M_LN2;
        }
        for (int j = 0; j < enc_n; ++j) {
            const double epsilon = errors[j];
            mismatch[i * enc_n + j] = log(gamma_xy_1m * (1 - epsilon) * n + gamma_xy * epsilon * (n / (n - 1))) / M_LN2;
        }
    }
}

static void aligner_free(aligner* A) {
    free(A->match[0]);
    free(A->mismatch[0]);
    free(A->scores);
    free(A->left_jump_scores);
    free(A->left_jump_points);
    free(A->directions);
}

static void aligner_reserve(aligner* A, size_t nrows) {
    if (nrows > A->cap_rows) {
        A->cap_rows = nrows;
        A->scores = (double*)realloc(A->scores, nrows * sizeof(double));
        A->left_jump_scores = (double*)realloc(A->left_jump_scores, nrows * sizeof(double));
        A->left_jump_points = (size_t*)realloc(A->left_jump_points, nrows * sizeof(size_t));
        A->directions = (int*)realloc(A->directions, nrows * (A->rlen + 1) * sizeof(int));
    }
}

/* src/reference_align.cpp:7-13 : note gap_open = go + ge */
static const char* aligner_init(aligner* A, const char* rseq, int enc_n, const char* const* enc_names,
        const double* enc_err, double go, double ge)
{
    memset(A, 0, sizeof(*A));
    A->rlen = strlen(rseq);
    A->rseq = rseq;
    A->gap_open = go + ge;
    A->gap_ext = ge;
    const char* msg = check_encoding(enc_n, enc_names, enc_err, &A->offset);
    if (msg) return msg;
    A->available = (size_t)enc_n;
    A->match[0] = (double*)malloc(4 * (size_t)enc_n * sizeof(double));
    A->mismatch[0] = (double*)malloc(4 * (size_t)enc_n * sizeof(double));
    for (int i = 1; i < 4; ++i) {
        A->match[i] = A->match[0] + (size_t)i * enc_n;
        A->mismatch[i] = A->mismatch[0] + (size_t)i * enc_n;
    }
    fill_tables(enc_n, enc_err, A->match[0], A->mismatch[0]);
    aligner_reserve(A, 1000);
    return NULL;
}

/* src/reference_align.cpp:214-225 */
static double precomputed_cost(aligner* A, int mode, int matched, char qual) {
    if (qual < A->offset) {
        A->error = "quality cannot be lower than smallest encoded value";
        return 0;
    }
    size_t location = (size_t)(qual - A->offset);
    if (location >= A->available) {
        location = A->available - 1;
    }
    return (matched ? A->match : A->mismatch)[mode - 1][location];
}

/* src/reference_align.cpp:184-212 : the IUPAC branches test `ref`, not `obs` (kept as is) */
static double compute_cost(aligner* A, char ref, char obs, char qual) {
    switch (ref) {
        case 'A': case 'C': case 'G': case 'T':
            return precomputed_cost(A, 1, (ref == obs), qual);
        case 'M': return precomputed_cost(A, 2, (ref == 'A' || ref == 'C'), qual);
        case 'R': return precomputed_cost(A, 2, (ref == 'A' || ref == 'G'), qual);
        case 'W': return precomputed_cost(A, 2, (ref == 'A' || ref == 'T'), qual);
        case 'S': return precomputed_cost(A, 2, (ref == 'C' || ref == 'G'), qual);
        case 'Y': return precomputed_cost(A, 2, (ref == 'C' || ref == 'T'), qual);
        case 'K': return precomputed_cost(A, 2, (ref == 'G' || ref == 'T'), qual);
        case 'V': return precomputed_cost(A, 3, (ref != 'T'), qual);
        case 'H': return precomputed_cost(A, 3, (ref != 'G'), qual);
        case 'D': return precomputed_cost(A, 3, (ref != 'C'), qual);
        case 'B': return precomputed_cost(A, 3, (ref != 'A'), qual);
        case 'N': return precomputed_cost(A, 4, 1, qual);
    }
    A->error = "unrecognized base in reference sequence";
    return 0;
}

/* src/reference_align.cpp:107-181 */
static void align_column(aligner* A, int* last_direction, size_t pos, size_t len, const char* seq, const char* qual, int last) {
    const char reference = A->rseq[pos];
    int* current_direction = last_direction + A->nrows;
    double* scores = A->scores;
    const double gap_open = A->gap_open, gap_ext = A->gap_ext;

    double lagging_last = scores[0];
    scores[0] -= (*last_direction > 0 ? gap_ext : gap_open);
    *current_direction = 1;

    double vert_gap_open = (last ? 0 : gap_open);
    double vert_gap_ext = (last ? 0 : gap_ext);
    double up_jump_score = -INFINITY;
    size_t up_jump_point = 0;

    for (size_t i = 1; i <= len; ++i) {
        double horiz_gap = scores[i] - (last_direction[i] > 0 ? gap_ext : gap_open);

        double* previous_horiz_gap = &A->left_jump_scores[i];
        *previous_horiz_gap -= gap_ext;
        size_t left_step = 1;
        if (*previous_horiz_gap > horiz_gap) {
            left_step = 1 + pos - A->left_jump_points[i];
            horiz_gap = *previous_horiz_gap;
        } else {
            *previous_horiz_gap = horiz_gap;
            A->left_jump_points[i] = pos;
        }

        double vert_gap = scores[i - 1] - (current_direction[i - 1] < 0 ? vert_gap_ext : vert_gap_open);

        up_jump_score -= vert_gap_ext;
        size_t up_step = 1;
        if (up_jump_score > vert_gap) {
            up_step = 1 + i - up_jump_point;
            vert_gap = up_jump_score;
        } else {
            up_jump_score = vert_gap;
            up_jump_point = i;
        }

        double match = lagging_last + compute_cost(A, reference, seq[i - 1], qual[i - 1]);
        if (A->error) return; /* the reference throws here */
        lagging_last = scores[i];

        if (match > horiz_gap && match > vert_gap) {
            current_direction[i] = 0;
            scores[i] = match;
        } else if (horiz_gap > vert_gap) {
            scores[i] = horiz_gap;
            current_direction[i] = (int)left_step;
        } else {
            scores[i] = vert_gap;
            current_direction[i] = -(int)up_step;
        }
    }
}

/* src/reference_align.cpp:54-105 */
static double aligner_align(aligner* A, size_t len, const char* seq, const char* qual, int local) {
    A->nrows = len + 1;
    aligner_reserve(A, A->nrows);
    const size_t nrows = A->nrows;

    for (size_t i = 0; i < nrows; ++i) A->directions[i] = -1;
    if (local) {
        for (size_t i = 0; i < nrows; ++i) A->scores[i] = 0;
    } else {
        A->scores[0] = 0;
        for (size_t i = 1; i < nrows; ++i) {
            A->scores[i] = -A->gap_open - A->gap_ext * (i - 1);
        }
    }
    for (size_t i = 0; i < nrows; ++i) {
        A->left_jump_scores[i] = -INFINITY;
        A->left_jump_points[i] = 0;
    }

    int* last_dir = A->directions;
    for (size_t col = 1; col < A->rlen; ++col) {
        align_column(A, last_dir, col - 1, len, seq, qual, 0);
        if (A->error) return 0;
        last_dir += nrows;
    }
    if (A->rlen) {
        align_column(A, last_dir, A->rlen - 1, len, seq, qual, local);
        if (A->error) return 0;
    }
    return A->scores[len];
}

/* Backtrack (src/reference_align.cpp:231-278) with the fill_map visitor (:280-305) folded in:
 * map_kind[c] = 1 if adaptor column c is a (mis)match at DP row map_row[c], 0 if it is a deletion
 * passed at DP row map_row[c]-1.  If rstr/qstr are non-NULL the fill_strings visitor (:353-390) runs too
 * (strings are produced reversed, then flipped by the caller). */
static size_t backtrack(const aligner* A, int* map_kind, size_t* map_row, char* rstr, char* qstr, const char* qseq) {
    const size_t nrows = A->nrows;
    const int* location = A->directions + nrows * A->rlen;
    size_t currow = nrows - 1;
    size_t nout = 0;

    for (size_t i = A->rlen; i > 0; --i) {
        while (currow > 0) {
            int curdir = location[currow];
            if (curdir >= 0) break;
            while (curdir < 0) {
                if (rstr) { rstr[nout] = '-'; qstr[nout] = qseq[currow - 1]; ++nout; }
                --currow;
                ++curdir;
            }
        }
        int curdir = location[currow];
        if (curdir == 0) {
            if (map_kind) { map_kind[i] = 1; map_row[i] = currow; }
            if (rstr) { rstr[nout] = A->rseq[i - 1]; qstr[nout] = qseq[currow - 1]; ++nout; }
            --currow;
            location -= nrows;
        } else {
            if (map_kind) { map_kind[i] = 0; map_row[i] = currow + 1; }
            if (rstr) { rstr[nout] = A->rseq[i - 1]; qstr[nout] = '-'; ++nout; }
            location -= nrows;
            while ((--curdir) > 0) {
                --i;
                if (map_kind) { map_kind[i] = 0; map_row[i] = currow + 1; }
                if (rstr) { rstr[nout] = A->rseq[i - 1]; qstr[nout] = '-'; ++nout; }
                location -= nrows;
            }
        }
    }
    while (currow > 0) {
        if (rstr) { rstr[nout] = '-'; qstr[nout] = qseq[currow - 1]; ++nout; }
        --currow;
    }
    return nout;
}

/* querymap::operator() (src/reference_align.cpp:307-351); map arrays have rlen+1 entries */
static void querymap_range(size_t rlen, size_t nrows, const int* map_kind, const size_t* map_row,
        size_t ref_start, size_t ref_end, int include_gaps, size_t* first, size_t* second)
{
    if (rlen + 1 <= 1) {
        *first = 0; *second = 0;
        return;
    }
    if (!include_gaps) {
        size_t curstart = map_row[ref_start + 1];
        size_t curend = map_row[ref_end];
        if (map_kind[ref_end]) ++curend;
        *first = curstart - 1; *second = curend - 1;
    } else {
        size_t curstart, curend;
        if (ref_start == 0) {
            curstart = 1;
        } else {
            curstart = map_row[ref_start];
            if (map_kind[ref_start]) ++curstart;
        }
        ++ref_end;
        if (ref_end == rlen + 1) {
            curend = nrows;
        } else {
            curend = map_row[ref_end];
        }
        *first = curstart - 1; *second = curend - 1;
    }
}

/* ---------------------------------------------------------------------------------------------- */

typedef struct {
    int kind; /* 0 adaptor_align, 1 score only, 2 general */
    int64_t n, lo, hi;
    const char *seq, *qual;
    const int64_t *seq_off, *qual_off;
    int enc_n; const char* const* enc_names; const double* enc_err;
    double go, ge;
    const char* reference;
    int local, edit_only;
    int nsec; const int32_t *sec_starts, *sec_ends;
    double* score; int32_t *start, *end, *sec_start, *sec_width, *edit;
    char *ref_aln, *query_aln; int64_t aln_stride;
    int64_t err_at; const char* err_msg;
} job;

static void* run_job(void* arg) {
    job* J = (job*)arg;
    aligner A;
    J->err_at = -1;
    const char* msg = aligner_init(&A, J->reference, J->enc_n, J->enc_names, J->enc_err, J->go, J->ge);
    if (msg) { J->err_at = 0; J->err_msg = msg; aligner_free(&A); return NULL; }
    const size_t rlen = A.rlen;
    int* map_kind = (int*)calloc(rlen + 1, sizeof(int));
    size_t* map_row = (size_t*)calloc(rlen + 1, sizeof(size_t));
    char *rwork = NULL, *qwork = NULL;
    size_t wcap = 0;

    for (int64_t i = J->lo; i < J->hi; ++i) {
        const char* sstr = J->seq + J->seq_off[i];
        const size_t slen = (size_t)(J->seq_off[i + 1] - J->seq_off[i]);
        if (slen != (size_t)(J->qual_off[i + 1] - J->qual_off[i])) {
            J->err_at = i; J->err_msg = "sequence and quality strings should have the same length";
            break;
        }
        const char* qstr = J->qual + J->qual_off[i];
        const int local = (J->kind == 0) ? 1 : (J->kind == 2 ? 0 : J->local);
        double sc = aligner_align(&A, slen, sstr, qstr, local);
        if (A.error) { J->err_at = i; J->err_msg = A.error; break; }
        J->score[i] = sc;

        if (J->kind == 0) {
            /* src/adaptor_align.cpp:54-68 */
            backtrack(&A, map_kind, map_row, NULL, NULL, NULL);
            size_t first, second;
            querymap_range(rlen, A.nrows, map_kind, map_row, 0, rlen, 0, &first, &second);
            if (first < second) {
                J->start[i] = (int32_t)(first + 1);
                J->end[i] = (int32_t)second;
            }
            for (int sec = 0; sec < J->nsec; ++sec) {
                querymap_range(rlen, A.nrows, map_kind, map_row, (size_t)J->sec_starts[sec], (size_t)J->sec_ends[sec], 1, &first, &second);
                J->sec_start[(int64_t)sec * J->n + i] = (int32_t)(first + 1);
                J->sec_width[(int64_t)sec * J->n + i] = (int32_t)(second - first);
            }
        } else if (J->kind == 2) {
            /* src/general_align.cpp:44-57 */
            if (slen + rlen + 1 > wcap) {
                wcap = slen + rlen + 1;
                rwork = (char*)realloc(rwork, wcap);
                qwork = (char*)realloc(qwork, wcap);
            }
            size_t nout = backtrack(&A, NULL, NULL, rwork, qwork, sstr);
            int32_t ed = 0;
            for (size_t j = 0; j < nout; ++j) {
                if (rwork[j] != qwork[j]) ++ed;
            }
            J->edit[i] = ed;
            if (!J->edit_only) {
                if ((int64_t)nout + 1 > J->aln_stride) { J->err_at = i; J->err_msg = "oracle: aln_stride too small"; break; }
                char* ro = J->ref_aln + i * J->aln_stride;
                char* qo = J->query_aln + i * J->aln_stride;
                for (size_t j = 0; j < nout; ++j) {
                    ro[j] = rwork[nout - 1 - j];
                    qo[j] = qwork[nout - 1 - j];
                }
                ro[nout] = '\0';
                qo[nout] = '\0';
            }
        }
    }
    free(map_kind); free(map_row); free(rwork); free(qwork);
    aligner_free(&A);
    return NULL;
}

static int run_all(job* proto, int nthreads, char* err, int errlen) {
    const int64_t n = proto->n;
    if (nthreads < 1) nthreads = 1;
    if ((int64_t)nthreads > n) nthreads = n > 0 ? (int)n : 1;
    job* jobs = (job*)malloc(sizeof(job) * (size_t)nthreads);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
    for (int t = 0; t < nthreads; ++t) {
        jobs[t] = *proto;
        jobs[t].lo = n * t / nthreads;
        jobs[t].hi = n * (t + 1) / nthreads;
    }
    if (nthreads == 1) {
        run_job(&jobs[0]);
    } else {
        for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, run_job, &jobs[t]);
        for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    }
    int rc = 0;
    for (int t = 0; t < nthreads; ++t) {
        if (jobs[t].err_at >= 0) {
            if (err && errlen > 0) { strncpy(err, jobs[t].err_msg, (size_t)errlen - 1); err[errlen - 1] = '\0'; }
            rc = 1;
            break;
        }
    }
    free(jobs); free(th);
    return rc;
}

int orc_adaptor_align(int64_t n, const char* seq, const int64_t* seq_off, const char* qual, const int64_t* qual_off,
        int enc_n, const char* const* enc_names, const double* enc_err, double go, double ge,
        const char* adaptor, int nsec, const int32_t* sec_starts, const int32_t* sec_ends,
        double* score, int32_t* start, int32_t* end, int32_t* sec_start, int32_t* sec_width,
        int nthreads, char* err, int errlen)
{
    job J; memset(&J, 0, sizeof(J));
    J.kind = 0; J.n = n; J.seq = seq; J.seq_off = seq_off; J.qual = qual; J.qual_off = qual_off;
    J.enc_n = enc_n; J.enc_names = enc_names; J.enc_err = enc_err; J.go = go; J.ge = ge; J.reference = adaptor;
    J.nsec = nsec; J.sec_starts = sec_starts; J.sec_ends = sec_ends;
    J.score = score; J.start = start; J.end = end; J.sec_start = sec_start; J.sec_width = sec_width;
    for (int64_t i = 0; i < n; ++i) { start[i] = 0; end[i] = 0; }
    return run_all(&J, nthreads, err, errlen);
}

int orc_align_score_only(int64_t n, const char* seq, const int64_t* seq_off, const char* qual, const int64_t* qual_off,
        int enc_n, const char* const* enc_names, const double* enc_err, double go, double ge,
        const char* reference, int local, double* score, int nthreads, char* err, int errlen)
{
    job J; memset(&J, 0, sizeof(J));
    J.kind = 1; J.n = n; J.seq = seq; J.seq_off = seq_off; J.qual = qual; J.qual_off = qual_off;
    J.enc_n = enc_n; J.enc_names = enc_names; J.enc_err = enc_err; J.go = go; J.ge = ge; J.reference = reference;
    J.local = local; J.score = score;
    return run_all(&J, nthreads, err, errlen);
}

int orc_general_align(int64_t n, const char* seq, const int64_t* seq_off, const char* qual, const int64_t* qual_off,
        int enc_n, const char* const* enc_names, const double* enc_err, double go, double ge,
        const char* reference, int edit_only, double* score, int32_t* edit,
        char* ref_aln, char* query_aln, int64_t aln_stride, int nthreads, char* err, int errlen)
{
    job J; memset(&J, 0, sizeof(J));
    J.kind = 2; J.n = n; J.seq = seq; J.seq_off = seq_off; J.qual = qual; J.qual_off = qual_off;
    J.enc_n = enc_n; J.enc_names = enc_names; J.enc_err = enc_err; J.go = go; J.ge = ge; J.reference = reference;
    J.edit_only = edit_only; J.score = score; J.edit = edit;
    J.ref_aln = ref_aln; J.query_aln = query_aln; J.aln_stride = aln_stride;
    for (int64_t i = 0; i < n; ++i) edit[i] = 0;
    return run_all(&J, nthreads, err, errlen);
}

int orc_cost_tables(int enc_n, const char* const* enc_names, const double* enc_err,
        double* match, double* mismatch, char* offset, char* err, int errlen)
{
    const char* msg = check_encoding(enc_n, enc_names, enc_err, offset);
    if (msg) {
        if (err && errlen > 0) { strncpy(err, msg, (size_t)errlen - 1); err[errlen - 1] = '\0'; }
        return 1;
    }
    fill_tables(enc_n, enc_err, match, mismatch);
    return 0;
}
