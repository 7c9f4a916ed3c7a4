/* sarlacc_b200 -- C ABI of the B200-native adaptor-alignment hot path.
 *
 * This is the drop-in boundary for the four `.Call` entry points of the reference R package
 * (registered in /root/reference/src/init.cpp:10-13, prototypes src/sarlacc.h:14-17):
 *
 *     reference `.Call` symbol (file:line)                     replaced by
 *     -------------------------------------------------------  ----------------------------------
 *     adaptor_align            src/adaptor_align.cpp:11-77      sarlacc_adaptor_align
 *     adaptor_align_score_only src/adaptor_align.cpp:79-110     sarlacc_adaptor_align_score_only
 *     barcode_align            src/barcode_align.cpp:10-44      sarlacc_barcode_align
 *     general_align            src/general_align.cpp:10-62      sarlacc_general_align
 *
 *     umi_group                src/umi_group.cpp:14-117         sarlacc_umi_group          (init.cpp:23)
 *     cluster_umis_test        src/cluster_umis_test.cpp:8-29   sarlacc_cluster_umis       (init.cpp:25)
 *     fast_levdist_test        src/sorted_trie.cpp:302-332      sarlacc_umi_neighbors      (init.cpp:24)
 *
 * plus fused entries that fold R-level loops into one device pass -- sarlacc_barcode_align_multi (the per-barcode
 * loop of R/barcodeAlign.R:20-35), sarlacc_adaptor_align_windows / sarlacc_adaptor_align_reads (.align_AA_internal,
 * R/adaptorAlign.R:178-199) --, a "resident" variant of the same calls that keeps packed read windows in HBM between
 * calls (what adaptorAlign -> getAdaptorThresholds -> tuneAlignment re-use), device-resident chunks re-loaded in place
 * (sarlacc_chunk_*: .align_AA_internal / .align_AT_internal on one FastqStreamer yield, R/adaptorAlign.R:26-48,
 * R/getAdaptorThresholds.R:35-48), threshold selection on the device (sarlacc_compute_threshold, sarlacc_tied_overlap),
 * and host-side helpers: FASTQ ingest (sarlacc_fastq_*, standing in for ShortRead::FastqStreamer, R/adaptorAlign.R:26,36),
 * the packer test hooks (sarlacc_pack_rows, sarlacc_pack_bases) and counters for bench.py (sarlacc_kernel_launches,
 * sarlacc_last_pair_timing, sarlacc_last_pair_upload_bytes).
 *
 * Plain pointers and sizes only: no R, Rcpp, torch or CUDA types appear in any signature.  The R-side
 * glue a maintainer would add (SEXP unpacking -> these calls) is shown in INTEGRATION.md and kept as
 * source in sarlacc_b200/csrc/r_glue.cpp.
 *
 * Conventions
 *   - every function returns 0 on success and non-zero on failure; the message (the reference's own
 *     std::runtime_error text where one exists) is then available from sarlacc_last_error() on the
 *     calling thread -- the glue passes it to Rf_error(), which is what BEGIN_RCPP/END_RCPP did.
 *   - inputs are caller-owned and never modified; outputs are caller-allocated.
 *   - there is NO CPU implementation behind these calls: without a usable CUDA device they fail with
 *     an error, they never fall back.
 */
#ifndef SARLACC_B200_H
#define SARLACC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SARLACC_SEQ_ASCII      0 /* character vector / already decoded: bytes are 'A','C','G','T',...   */
#define SARLACC_SEQ_BIOSTRINGS 1 /* DNAStringSet payload: Biostrings DNA byte codes (A=1,C=2,G=4,T=8,...)
                                    i.e. what src/DNA_input.cpp:64-75 runs DNAdecode() over               */

/* A set of n (sequence, quality) string pairs.  Two layouts are accepted:
 *   views : seq[i]/qual[i] point at the i-th string (what get_elt_from_XStringSet_holder returns for
 *           every element, src/adaptor_align.cpp:46-50); seq_len[i]/qual_len[i] are the lengths.
 *   CSR   : seq == NULL; string i is seq_pool[seq_off[i] .. seq_off[i+1]) (same for qual).
 * Sequence and quality lengths are passed separately so that the reference's "sequence and quality
 * strings should have the same length" check (src/adaptor_align.cpp:51-53) is made here, not upstream. */
typedef struct {
    int64_t n;
    int seq_encoding;               /* SARLACC_SEQ_ASCII or SARLACC_SEQ_BIOSTRINGS */
    const uint8_t* const* seq;      /* views layout (or NULL)  */
    const int32_t* seq_len;
    const uint8_t* const* qual;
    const int32_t* qual_len;
    const uint8_t* seq_pool;        /* CSR layout (used when seq == NULL) */
    const int64_t* seq_off;
    const uint8_t* qual_pool;
    const int64_t* qual_off;
} sarlacc_reads;

/* The named numeric vector every hot-path .Call receives from .create_encoding_vector
 * (R/qualityMask.R:19-27): names[i] is the i-th name as a C string, err[i] its error probability.
 * Validation follows quality_encoding's constructor (src/quality_encoding.cpp:5-32) message for message. */
typedef struct {
    int n;
    const char* const* names;       /* NULL = unnamed vector */
    const double* err;
} sarlacc_encoding;

/* ---- process-wide state ------------------------------------------------------------------------ */
const char* sarlacc_last_error(void);           /* thread-local, valid until the next call on this thread */
int  sarlacc_device_count(void);                /* CUDA devices visible; <0 on error */
int  sarlacc_set_devices(const int* devices, int ndevices); /* devices the host-buffer calls shard reads over
                                                   (contiguous read-index ranges, R/adaptorAlign.R:126-134).
                                                   Default: device 0 only. */
int  sarlacc_set_host_threads(int nthreads);    /* packer threads per device (default: hardware concurrency / devices) */
const char* sarlacc_version(void);
/* Launch accounting for bench.py's "gpu_launches": kernels launched by this library since the last reset. */
int64_t sarlacc_kernel_launches(int reset);
/* Device buffers released by the library are kept for reuse (SARLACC_POOL_MB, default 24576; 0 = off); this hands them
 * back to the driver. */
void sarlacc_trim_device_memory(void);
/* Phases of the last sarlacc_adaptor_align_windows / _reads call (first device), milliseconds: host side -- staging and
 * length scans, enqueueing, waiting for results + copy-out, total --, then two device-side sums over the call's chunks
 * from CUDA events: upload (H2D + device packer) and kernels + copy back (chunks overlap, so these exceed the wall time).
 * For bench.py's per-rank breakdown. */
void sarlacc_last_pair_timing(double* ms6);
/* Bytes of window data the same call copied to the (first) device: bases + qualities + offsets of both window sets.  With
 * SARLACC_PACK_SEQ=1 in the environment the bases go up as 4-bit codes -- one pass of the host over the sequence bytes for
 * 25 % fewer bytes on the link; worth it only where host cores are plentiful and the link is not (off by default). */
int64_t sarlacc_last_pair_upload_bytes(void);

/* ---- the four reference entry points (host buffers in, host buffers out) ------------------------- */

/* adaptor_align(readseq, readqual, encoding, gapopen, gapext, adaptor, sec_starts, sec_ends)
 * src/adaptor_align.cpp:11-77.  Local-in-read / global-in-adaptor alignment with traceback.
 *   sec_starts are 0-based, sec_ends 1-based, as R passes them (R/adaptorAlign.R:158).
 *   score[n]; start[n], end[n] are 1-based read coordinates or 0,0 when the guard at :58 fails;
 *   sec_start / sec_width are [nsec][n] (section-major, one R IntegerVector per section, :64-68). */
int sarlacc_adaptor_align(const sarlacc_reads* reads, const sarlacc_encoding* encoding,
        double gapopen, double gapext, const char* adaptor,
        int nsec, const int32_t* sec_starts, const int32_t* sec_ends,
        double* score, int32_t* start, int32_t* end, int32_t* sec_start, int32_t* sec_width);

/* adaptor_align_score_only(readseq, readqual, encoding, gapopen, gapext, adaptor)  src/adaptor_align.cpp:79-110 */
int sarlacc_adaptor_align_score_only(const sarlacc_reads* reads, const sarlacc_encoding* encoding,
        double gapopen, double gapext, const char* adaptor, double* score);

/* barcode_align(barcodeseq, barcodequal, encoding, gapopen, gapext, reference)  src/barcode_align.cpp:10-44
 * Fully global alignment, score only. */
int sarlacc_barcode_align(const sarlacc_reads* reads, const sarlacc_encoding* encoding,
        double gapopen, double gapext, const char* reference, double* score);

/* general_align(inputseq, inputqual, encoding, gapopen, gapext, reference, edit_only)  src/general_align.cpp:10-62
 * Global alignment + gapped strings + edit distance.  When edit_only == 0, ref_aln/query_aln receive
 * NUL-terminated strings at i*aln_stride (aln_stride >= max len + strlen(reference) + 1). */
int sarlacc_general_align(const sarlacc_reads* reads, const sarlacc_encoding* encoding,
        double gapopen, double gapext, const char* reference, int edit_only,
        double* score, int32_t* edit, char* ref_aln, char* query_aln, int64_t aln_stride);

/* ---- fused extension: every barcode in one pass ---------------------------------------------------
 * Replaces the body of the `for (b in seq_along(barcodes))` loop of R/barcodeAlign.R:20-35 (nbarcodes
 * calls of barcode_align + the running best / next-best update with its strict `>` rule):
 *   best_id[n]  1-based index of the best barcode (0 where R would hold NA, i.e. nbarcodes == 0 or all NaN)
 *   best[n]     its score (-Inf if none);  next_best[n] the runner-up (-Inf if none).
 * R's `gap` column is best - next_best (R/barcodeAlign.R:37).  all_scores may be NULL; otherwise it
 * receives the [nbarcodes][n] matrix of scores exactly as nbarcodes barcode_align calls would. */
int sarlacc_barcode_align_multi(const sarlacc_reads* reads, const sarlacc_encoding* encoding,
        double gapopen, double gapext, const char* const* barcodes, int nbarcodes,
        int32_t* best_id, double* best, double* next_best, double* all_scores);

/* ---- fused extension: both adaptors on both read ends ---------------------------------------------
 * Replaces the body of .align_AA_internal (R/adaptorAlign.R:178-199: four adaptor_align calls -- (adaptor1, front),
 * (adaptor2, back), (adaptor1, back), (adaptor2, front) --, .resolve_strand (:112-122) and the per-row selection
 * cur.starts[rev,] <- cur.rc.starts[rev,]) plus adaptorAlign's adaptor2 coordinate flip width - x + 1 (:66-71,
 * applied when read_width != NULL).  `front` / `back` are the windows .get_front_and_back (:86-95) produced (the
 * back one already reverse-complemented), element i of both belonging to read i.  Outputs are the rows R keeps:
 * reversed[n] (0/1), and for adaptor k the score / start / end / [nsec_k][n] section starts and widths of the
 * alignment on the selected window.  Results are identical to composing the four calls (tests/test_gpu_api.py). */
int sarlacc_adaptor_align_windows(const sarlacc_reads* front, const sarlacc_reads* back, const sarlacc_encoding* encoding,
        double gapopen, double gapext, const char* adaptor1, const char* adaptor2,
        int nsec1, const int32_t* sec_starts1, const int32_t* sec_ends1,
        int nsec2, const int32_t* sec_starts2, const int32_t* sec_ends2,
        const int32_t* read_width, uint8_t* reversed,
        double* score1, int32_t* start1, int32_t* end1, int32_t* sec_start1, int32_t* sec_width1,
        double* score2, int32_t* start2, int32_t* end2, int32_t* sec_start2, int32_t* sec_width2);

/* The same with WHOLE reads in and .get_front_and_back (R/adaptorAlign.R:86-95) done by the packer: front window =
 * first min(tolerance, width) bases, back window = reverse complement of the last min(tolerance, width) bases, cut
 * and complemented on the fly while packing (never materialised on the host).  read_width[n] receives width(reads);
 * adaptor2's start/end are always flipped into read coordinates (width - x + 1).  Needs two non-empty adaptors. */
int sarlacc_adaptor_align_reads(const sarlacc_reads* reads, int tolerance, const sarlacc_encoding* encoding,
        double gapopen, double gapext, const char* adaptor1, const char* adaptor2,
        int nsec1, const int32_t* sec_starts1, const int32_t* sec_ends1,
        int nsec2, const int32_t* sec_starts2, const int32_t* sec_ends2,
        int32_t* read_width, uint8_t* reversed,
        double* score1, int32_t* start1, int32_t* end1, int32_t* sec_start1, int32_t* sec_width1,
        double* score2, int32_t* start2, int32_t* end2, int32_t* sec_start2, int32_t* sec_width2);

/* ---- resident read windows -----------------------------------------------------------------------
 * Packs the reads once (2 bytes per base: quality index + one-hot base), uploads them and keeps them
 * in HBM.  The *_resident calls then run on that copy; results stay on the device until fetched.
 * `stream` is a cudaStream_t passed as void* (NULL = the library's own stream for that object); the
 * call only enqueues work, so the caller may bracket it with its own events. */
typedef struct sarlacc_resident sarlacc_resident;

sarlacc_resident* sarlacc_resident_create(const sarlacc_reads* reads, const sarlacc_encoding* encoding, int device);
/* Device-side .scramble_input (R/getAdaptorThresholds.R:68-92): a new resident object whose window i is a uniform
 * random permutation of window i of `src` (bases and qualities move together).  The permutation is a pure function of
 * (seed, read index, stream_id, length): read index = read_index[i] if given, else first_index + i -- so it does not
 * depend on chunking, sharding or device count, and sarlacc_b200/api.py reproduces it on the host for parity. */
sarlacc_resident* sarlacc_resident_scrambled(const sarlacc_resident* src, uint64_t seed, uint64_t first_index,
        const uint64_t* read_index, int stream_id);
/* Copies the packed rows (uint16[n][stride]: quality index | one-hot base << 8) and lengths back (tests). */
int     sarlacc_resident_rows(sarlacc_resident* r, uint16_t* rows, int32_t* lens, int* stride);
void    sarlacc_resident_free(sarlacc_resident* r);
int64_t sarlacc_resident_n(const sarlacc_resident* r);
int64_t sarlacc_resident_cells(const sarlacc_resident* r, int rlen);   /* sum(len_i) * rlen: DP cells of one pass */
int64_t sarlacc_resident_bytes(const sarlacc_resident* r);             /* packed bytes resident in HBM */

/* mode: 0 = score only, local (adaptor_align_score_only); 1 = local + traceback (adaptor_align);
 *       2 = score only, global (barcode_align). */
int sarlacc_resident_align(sarlacc_resident* r, int mode, double gapopen, double gapext, const char* reference,
        int nsec, const int32_t* sec_starts, const int32_t* sec_ends, void* stream);
/* Copies the results of the last sarlacc_resident_align to the host (any pointer may be NULL). */
int sarlacc_resident_fetch(sarlacc_resident* r, double* score, int32_t* start, int32_t* end,
        int32_t* sec_start, int32_t* sec_width, void* stream);
/* Device pointer of the last run's score vector (double[n]) -- for callers that keep working on the device. */
const double* sarlacc_resident_scores_device(const sarlacc_resident* r);
/* Name of the forward kernel variant the last sarlacc_resident_align used (for reports), e.g. "wf<C=9,G=8,trace>". */
const char* sarlacc_resident_last_kernel(const sarlacc_resident* r);
/* Roofline accounting: when enabled, every forward-kernel launch of sarlacc_resident_align is bracketed by CUDA
 * events on the launching stream; sarlacc_resident_forward_ms() then returns the summed device time of the
 * last run's forward launches (synchronises on them). */
void   sarlacc_resident_set_timing(sarlacc_resident* r, int on);
double sarlacc_resident_forward_ms(sarlacc_resident* r);

/* ---- chunks: device-resident reads, re-loaded in place (BASELINE.json configs[2] and configs[4]) -----
 * One chunk is to this library what one FastqStreamer yield() is to the R drivers (R/adaptorAlign.R:26-48,
 * R/getAdaptorThresholds.R:35-48): up to `capacity` reads whose front and back windows (.get_front_and_back,
 * R/adaptorAlign.R:86-95) sit packed in HBM.  Loading replaces the contents in place; the two passes below are the
 * bodies of the per-chunk workers:
 *     sarlacc_chunk_adaptor_align     .align_AA_internal (R/adaptorAlign.R:178-199) + adaptor2's coordinate flip (:66-71)
 *     sarlacc_chunk_scrambled_scores  .align_AT_internal (R/getAdaptorThresholds.R:105-128)
 * Every call only ENQUEUES work on the chunk's streams; outputs may be host (ideally page-locked) or device pointers and
 * are complete when sarlacc_chunk_sync returns, so the copy-out of one chunk overlaps the alignment of the next.  Output
 * pointers may be NULL.  Section matrices are [nsec][out_pitch] (out_pitch >= reads in the chunk: the caller may point
 * into the columns of one big result table).  Results are those of composing the four reference calls
 * (tests/test_gpu_chunk.py): only the strand .resolve_strand keeps is ever walked back, which the R code computes and
 * then discards. */
typedef struct sarlacc_chunk sarlacc_chunk;
sarlacc_chunk* sarlacc_chunk_create(int device, int64_t capacity, int tolerance, const sarlacc_encoding* encoding);
void    sarlacc_chunk_free(sarlacc_chunk* c);
int64_t sarlacc_chunk_n(const sarlacc_chunk* c);
/* Host reads in.  tolerance == 0: `front` / `back` are the windows as .get_front_and_back made them (back already
 * reverse-complemented) and `width` the read widths or NULL (then adaptor2's coordinates stay window-relative);
 * tolerance > 0: `front` holds WHOLE reads, `back` must be NULL, both windows are cut by the packer.  Returns the
 * reference's per-read errors ("sequence and quality strings should have the same length", "quality cannot be lower
 * than smallest encoded value") before anything is aligned. */
int sarlacc_chunk_load_reads(sarlacc_chunk* c, const sarlacc_reads* front, const sarlacc_reads* back, int tolerance, const int32_t* width);
/* Synthetic reads generated on the device: the mockReads recipe (R/mockReads.R:58-92) with a counter-based generator
 * keyed by (seed, first_index + i), so read i of a run is the same whatever the chunking, sharding or device count;
 * sarlacc_b200/synth.py: mock_windows is its host mirror, bit for bit.  barcodes may be NULL (the barcode slot -- adaptor1's
 * first N run -- then holds one random base repeated, :50).  Only the two windows and the read width are materialised. */
int sarlacc_chunk_load_mock(sarlacc_chunk* c, int64_t n, uint64_t first_index, uint64_t seed,
        const char* adaptor1, const char* adaptor2, int insert_len, const char* const* barcodes, int nbarcodes,
        double sub_rate, double indel_rate, int max_insert);
int sarlacc_chunk_adaptor_align(sarlacc_chunk* c, double gapopen, double gapext, const char* adaptor1, const char* adaptor2,
        int nsec1, const int32_t* sec_starts1, const int32_t* sec_ends1,
        int nsec2, const int32_t* sec_starts2, const int32_t* sec_ends2,
        int64_t out_pitch, int32_t* read_width, uint8_t* reversed,
        double* score1, int32_t* start1, int32_t* end1, int32_t* sec_start1, int32_t* sec_width1,
        double* score2, int32_t* start2, int32_t* end2, int32_t* sec_start2, int32_t* sec_width2);
/* scramble == 1: both windows of every read are permuted first (.scramble_input, R/getAdaptorThresholds.R:68-92; the
 * permutation of read i depends on (seed, read_index[i] or first_index + i) only -- sarlacc_b200/api.py: _scramble_by_index
 * is the host mirror); scramble == 2 re-uses the permuted windows of the previous call (tuneAlignment scores one
 * scramble under 35 penalty pairs, R/tuneAlignment.R:27-72); scramble == 0 scores the windows as loaded
 * (.get_alignment_scores, R/tuneAlignment.R:99-112).  score1 / score2 receive ifelse(is.reverse, revcomp score, forward
 * score) per adaptor (R/getAdaptorThresholds.R:123-127); strand_score receives .resolve_strand()$scores, the larger of
 * the two strand sums (R/adaptorAlign.R:112-122), which is what tuneAlignment compares. */
int sarlacc_chunk_scrambled_scores(sarlacc_chunk* c, double gapopen, double gapext, const char* adaptor1, const char* adaptor2,
        uint64_t seed, uint64_t first_index, const uint64_t* read_index, int scramble, double* score1, double* score2,
        double* strand_score);
int sarlacc_chunk_sync(sarlacc_chunk* c);
/* The chunk's compute stream (a cudaStream_t as void*), and a join that makes it wait for everything enqueued so far on
 * the chunk's traceback and copy streams: an event recorded on the stream after sarlacc_chunk_join marks the completion of
 * all prior calls (how bench.py brackets its timed region with CUDA events). */
void* sarlacc_chunk_stream(sarlacc_chunk* c);
int   sarlacc_chunk_join(sarlacc_chunk* c);
/* Tests / reports: packed rows (uint16[n][stride]) of window set `which` (0 front, 1 back, 2 scrambled front,
 * 3 scrambled back), window lengths, read widths, strand flips of the generator; any pointer may be NULL. */
int sarlacc_chunk_rows(sarlacc_chunk* c, int which, uint16_t* rows, int32_t* lens, int* stride, int32_t* width, uint8_t* flipped);
/* Phase accounting with CUDA events on the chunk's compute stream: ms4 = {load / generate, adaptor_align forward passes,
 * scramble, score-only passes} summed since timing was switched on. */
void sarlacc_chunk_set_timing(sarlacc_chunk* c, int on);
int  sarlacc_chunk_phase_ms(sarlacc_chunk* c, double* ms4);
const char* sarlacc_chunk_last_kernel(const sarlacc_chunk* c, int adaptor);   /* forward kernel of adaptor 0 / 1 in the last adaptor_align */

/* .compute_threshold (R/getAdaptorThresholds.R:94-103): both vectors sorted (device radix sort), fdr_k = (n_scr -
 * #{scrambled <= real_k}) / (n_real - k), threshold = real[min(which(fdr <= error))]; NaN where R returns NA.  The
 * vectors may live on the host or on `device` (e.g. gathered from all ranks); they are not modified. */
int sarlacc_compute_threshold(const double* real, int64_t nreal, const double* scrambled, int64_t nscr, double error,
                              int device, double* threshold);
/* .tied_overlap (R/tuneAlignment.R:78-86): sum((findInterval(real, fake) + findInterval(real, fake, left.open=TRUE)) / 2)
 * / (length(real) * length(fake)) with fake sorted on the device; vectors on the host or on `device`. */
int sarlacc_tied_overlap(const double* real, int64_t nreal, const double* fake, int64_t nfake, int device, double* overlap);

/* ---- UMI grouping ----------------------------------------------------------------
 * Replaces SEXP umi_group(umi1, thresh1, umi2, thresh2, pregroup) (src/umi_group.cpp:14-117, registered at
 * src/init.cpp:23) together with unlist(out, recursive=FALSE) of R/umiGroup.R:22: the bounded masked-Levenshtein
 * neighbour search of src/sorted_trie.cpp runs as an all-pairs pass on the device, the greedy clustering of
 * src/cluster_umis.cpp on the host.  UMIs are ASCII (the decoded form process_DNA_input yields, src/DNA_input.cpp:64-88)
 * as one pool + n+1 offsets; umi2_pool may be NULL (one UMI).  Pre-groups are given like R's by.group list: ngroups+1
 * offsets into 1-based read indices.  The result is a list of integer vectors (1-based read indices per cluster, in
 * the reference's order), held by a handle:
 *     count  = number of vectors,  values = total length;  fetch copies count+1 offsets and the values.
 * Returns NULL with sarlacc_last_error() set on failure; the reference's messages are kept ("single-read groups
 * should contain only the read itself", "zero length read group", "'umi1' and 'umi2' should have the same length").
 * sarlacc_umi_neighbors returns the neighbour lists themselves (one vector per read of every pre-group, in group
 * order) -- with a single pre-group 1..n that is fast_levdist_test(seqs, limit, TRUE) (src/sorted_trie.cpp:302-332). */
typedef struct sarlacc_lists sarlacc_lists;
sarlacc_lists* sarlacc_umi_group(const uint8_t* umi1_pool, const int64_t* umi1_off, int64_t n, int threshold1,
                                 const uint8_t* umi2_pool, const int64_t* umi2_off, int threshold2,
                                 const int64_t* group_off, const int32_t* group_members, int64_t ngroups, int device);
sarlacc_lists* sarlacc_umi_neighbors(const uint8_t* umi1_pool, const int64_t* umi1_off, int64_t n, int threshold1,
                                     const uint8_t* umi2_pool, const int64_t* umi2_off, int threshold2,
                                     const int64_t* group_off, const int32_t* group_members, int64_t ngroups, int device);
/* The clustering step alone, on the host -- SEXP cluster_umis_test(links) (src/cluster_umis_test.cpp:8-29, registered at
 * src/init.cpp:25): n lists of 1-based neighbour indices (n+1 offsets), 1-based clusters out.  No device involved. */
sarlacc_lists* sarlacc_cluster_umis(const int64_t* link_off, const int32_t* links, int64_t n);
int64_t sarlacc_lists_count(const sarlacc_lists* r);
int64_t sarlacc_lists_values(const sarlacc_lists* r);
int     sarlacc_lists_fetch(const sarlacc_lists* r, int64_t* off, int32_t* values);
void    sarlacc_lists_free(sarlacc_lists* r);

/* ---- host packer, exposed for tests (no device involved) ------------------------------------------
 * Packs reads [0, n) the way every entry point above does before upload: rows[i*stride + r] = quality index
 * (min(qual - offset, |enc| - 1), src/reference_align.cpp:218-221) | one-hot base << 8 (src/DNA_input.cpp:64-75
 * folded in).  tolerance > 0 cuts the window .get_front_and_back would (R/adaptorAlign.R:86-95): the first
 * min(tolerance, width) bases, or with back != 0 the reverse complement of the last ones.  force_scalar != 0 bypasses
 * the vector packer.  lens[i] receives the window length (0 for a sequence/quality length mismatch); stride must be
 * >= the longest window.  Returns 0, or 1 with sarlacc_last_error() set (the reference's messages). */
int sarlacc_pack_rows(const sarlacc_reads* reads, const sarlacc_encoding* encoding, int tolerance, int back,
                      int stride, uint16_t* rows, int32_t* lens, int force_scalar);

/* The host half of the optional 4-bit upload (SARLACC_PACK_SEQ=1): out[(n + 1) / 2] <- one-hot base codes of seq[0, n), two
 * per byte (base k in nibble k & 1 of byte k >> 1; A = 1, C = 2, G = 4, T = 8, anything else 0 -- the table the device
 * packer applies to plain bytes, src/DNA_input.cpp:64-75 for Biostrings codes).  force_scalar: the table loop instead of
 * the AVX2 one.  Exposed for tests. */
int sarlacc_pack_bases(const uint8_t* seq, int64_t n, int seq_encoding, uint8_t* out, int force_scalar);

/* ---- FASTQ ingest (host side; stands in for ShortRead::FastqStreamer + .FASTQ2QSDS, R/adaptorAlign.R:26,36,104-110) --
 * Buffered reader of plain-text 4-line FASTQ records yielding chunks as CSR pools that can be handed straight back as a
 * sarlacc_reads (CSR layout).  Names exclude the leading '@'.  Pointers stay valid until the next call on the handle. */
typedef struct sarlacc_fastq sarlacc_fastq;
sarlacc_fastq* sarlacc_fastq_open(const char* path);
int64_t sarlacc_fastq_next(sarlacc_fastq* f, int64_t max_reads,
        const uint8_t** seq_pool, const int64_t** seq_off, const uint8_t** qual_pool, const int64_t** qual_off,
        const uint8_t** name_pool, const int64_t** name_off);   /* reads returned; 0 at end of file; -1 on a malformed record */
void sarlacc_fastq_close(sarlacc_fastq* f);
/* Parallel, condensed variant for the adaptor path: the file is mapped and parsed by `nthreads` threads (0 = auto), and
 * of every read only its name, its length (`width`) and its first and last `keep` bases + qualities are kept (the whole
 * read when it is no longer than 2*keep) -- all .get_front_and_back (R/adaptorAlign.R:86-95) ever looks at for
 * tolerance <= keep.  The condensed reads come back as CSR pools like sarlacc_fastq_next's and can be handed to
 * sarlacc_adaptor_align_reads unchanged: their windows are the original read's windows; only adaptor2's coordinate flip
 * needs the true width (R/adaptorAlign.R:66-71).  Do not mix with sarlacc_fastq_next on one handle.  Records whose
 * sequence and quality lengths differ are rejected as malformed. */
int64_t sarlacc_fastq_next_condensed(sarlacc_fastq* f, int64_t max_reads, int keep, int nthreads,
        const uint8_t** seq_pool, const int64_t** seq_off, const uint8_t** qual_pool, const int64_t** qual_off,
        const uint8_t** name_pool, const int64_t** name_off, const int32_t** width);

#ifdef __cplusplus
}
#endif

#endif
