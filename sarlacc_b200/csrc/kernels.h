/* Internal interface between the host driver (api.cpp) and the CUDA kernels (kernels.cu).
 *
 * Data layout in HBM (see DESIGN.md "Data layout"):
 *   rows   uint16[n][stride]   one entry per read base: low byte = quality index min(qual-offset, |enc|-1)
 *                              (reference clamp: src/reference_align.cpp:218-221), high byte = one-hot base
 *                              A=1 C=2 G=4 T=8, 0 for anything else (never equal to an ACGT reference base,
 *                              which is all src/reference_align.cpp:186-187 ever tests).
 *   lens   int32[n]            min(tolerance, read length) etc.; 0 allowed (handled without a DP).
 *   flags  4 bits per DP cell  {pd, p5, p1, p2} = {diag strictly best, horiz > vert, E-extend raw, F-extend raw}
 *                              -- the compact equivalent of the reference's int direction matrix
 *                              (src/reference_align.cpp:164-174 + the jump lengths of :134-155), see traceback.
 */
#ifndef SARLACC_B200_KERNELS_H
#define SARLACC_B200_KERNELS_H

#include <cuda_runtime.h>
#include <stdint.h>

namespace sarlacc {

constexpr int kMaxC = 18;        /* adaptor columns per lane in the wavefront kernel (instantiated: 1..12, 14, 16, 18) */
constexpr int kSoloMinC = 20;    /* one-thread-per-alignment geometry ("solo", G = 1, C = L): instantiated for L = 20..24 */
constexpr int kSoloMaxC = 24;
constexpr int kMaxGroup = 32;    /* lanes per alignment */
#ifndef SARLACC_WF_SKEW
#define SARLACC_WF_SKEW 2
#endif
constexpr int kSkew = SARLACC_WF_SKEW;   /* rows by which lane j+1 lags lane j (see kernels.cu) */
constexpr int kMaxFastL = kMaxC * kMaxGroup;

/* Column classes of a reference position (src/reference_align.cpp:184-212). */
enum ColKind : uint8_t {
    COL_ACGT = 0,   /* table m=1, match iff obs == ref                       */
    COL_TWO = 1,    /* M R W S Y K -> mismatch table m=2 (quirk kept, :188-199) */
    COL_THREE = 2,  /* V H D B     -> match table m=3                        */
    COL_N = 3       /* N           -> match table m=4                        */
};

struct AlignArgs {
    /* reads */
    const uint16_t* rows;
    const int32_t* lens;
    long long n;
    int stride;
    /* optional processing order: the a-th alignment of the launch is read index[a] (rows, lens, scores and the running
     * best are addressed by read; traceback records and endrow by a).  Barcode-length reads are walked in order of length,
     * so that the lanes of a warp -- independent alignments -- reach their ends together (api.cpp: DeviceJob). */
    const int32_t* index;
    /* optional dynamic distribution (row-pair kernels): groups take the next position of the processing order from this
     * device counter (zeroed by the caller) instead of striding over it, so a launch has no tail round whatever its
     * length -- which may itself live on the device: range[0], range[1] = first position and end (null: 0, n). */
    unsigned long long* next;
    const int32_t* range;
    /* reference(s): nref strings of length L (nref > 1 only for the fused multi-barcode pass) */
    int L;
    int nref;
    const uint8_t* refmask;   /* [nref][L] one-hot nibble for ACGT columns, 0 otherwise */
    const uint8_t* refkind;   /* [nref][L] ColKind */
    /* scoring */
    int local;                /* 1: local in read / global in reference, last column has free vertical gaps */
    double gop, ge;           /* gap_open = go + ge, gap_ext = ge (src/reference_align.cpp:8) */
    const double* row0;       /* [L+1] H[0][c] (row 0 chain of src/reference_align.cpp:116-117) */
    const double* cost;       /* [5][enc_n]: match1, mismatch1, mismatch2, match3, match4 */
    int enc_n;
    int kinds;                /* bit k set if some reference column has ColKind k */
    /* wavefront geometry */
    int G, C;
    int pair_rows;            /* 1: wf_forward2 (two rows per lane step) */
    unsigned one;             /* the value 1, opaque to the compiler (kernels.cu: trace_bit) */
    /* outputs */
    double* score;            /* [nref][n] (may be null when nref > 1 and only the reduction is wanted) */
    int32_t* best_id;         /* [n] multi-reference running best (R/barcodeAlign.R:28-34), nref > 1 only */
    double* best;
    double* next_best;
    /* traceback records: 4 bits per cell.  Wavefront layout: word (row slot, lane) holds the lane's first 8 (C <= 8) or 16
     * columns in `flags` (32- / 64-bit words) and columns 16.. in `flags_hi` (one byte per word for C = 17, 18; one
     * 32-bit word in the solo geometry), both indexed alike -- no padding bits are written (C = 18: 9 bytes per lane-row). */
    void* flags;
    void* flags_hi;
    long long fstride;        /* per alignment, in flag words (wavefront), row slots (solo) or bytes (generic) */
    int solo;                 /* 1: one thread per alignment (G = 1, C = L), records laid out [32 alignments][row][lane] */
    int32_t* endrow;          /* [n] wavefront + trace: row where the traceback's climb up the last column lands (optional) */
    /* generic kernel scratch: [maxlen+1][nthreads] doubles each */
    double* gS;
    double* gE;
    uint8_t* gChoice;
    long long gthreads;
};

struct TraceArgs {
    const int32_t* lens;
    long long n;
    int L;
    int layout;               /* 0: wavefront words, 1: generic bytes, 2: solo planes */
    int G, C, wordbytes;      /* wordbytes: 4 or 8 (the low word) */
    int hi_bytes;             /* 0, 1 or 4: element size of flags_hi (columns 16.. of a lane) */
    const void* flags;
    const void* flags_hi;
    long long fstride;
    const int32_t* endrow;      /* [n] optional: start the walk at (endrow[a], L) instead of (len, L) */
    /* fused both-ends runs: a second source (the same adaptor on the other window set) and, per alignment, which of the
     * two .resolve_strand kept (0: lens/flags/endrow, 1: lens2/flags2/endrow2; sel == null: always the first) */
    const uint8_t* sel;
    const int32_t* lens2;
    const void* flags2;
    const void* flags_hi2;
    const int32_t* endrow2;
    /* optional: where the records of read r sit in each source (position in that forward launch's processing order);
     * null: position == read */
    const int32_t* pos;
    const int32_t* pos2;
    const int32_t* width;       /* optional read widths: start/end are flipped to width - x + 1 (R/adaptorAlign.R:66-71) */
    long long out_pitch;        /* row pitch of sec_start / sec_width (0: n) */
    int nsec;
    const int32_t* sec_starts;  /* device, 0-based */
    const int32_t* sec_ends;    /* device, 1-based */
    int32_t* map;               /* scratch [L+1][n] */
    int32_t* start;
    int32_t* end;
    int32_t* sec_start;         /* [nsec][n] */
    int32_t* sec_width;
    uint8_t* ops;               /* optional [n][ops_stride] alignment operations, written back to front */
    int32_t* nops;
    long long ops_stride;
};

/* .resolve_strand on four device score vectors (R/adaptorAlign.R:112-122): reversed (may be null) and the two kept scores. */
struct StrandArgs {
    long long n;
    const double* a1_front; const double* a2_back; const double* a1_back; const double* a2_front;
    uint8_t* reversed;
    double* score1; double* score2;       /* ifelse(is.reverse, revcomp, forward) per adaptor; may be null */
    double* strand_score;                 /* .resolve_strand()$scores = ifelse(is.reverse, rscore, fscore); may be null */
    /* speculative runs (see StrandLists): reads whose kept strand was scored without records are appended to that
     * strand's list behind the predicted ones, for a second forward pass with records */
    const uint8_t* predicted;             /* null: no speculation */
    int32_t* list_fwd; int32_t* list_rev; int32_t* pos_fwd; int32_t* pos_rev;
    int32_t* ranges;
};
void launch_resolve_strand(const StrandArgs& s, cudaStream_t st);

/* Speculative traceback records.  Only the strand .resolve_strand keeps is walked back, but which one that is is known
 * only after all four forward passes -- so all four used to write records.  A cheap predictor (exact 8-mer seeds of the
 * two adaptors counted in both windows) sorts the reads of a launch into "forward strand", "reverse strand" and "unsure":
 * the passes of the predicted strand run with records, those of the other strand score-only (unsure reads: both with
 * records), and the few reads whose prediction .resolve_strand contradicts get a second forward pass with records.  The
 * results are those of the unspeculated run -- the predictor only decides where records are written.
 *   lists: list_fwd / list_rev = reads whose forward- / reverse-strand passes write records (predicted ones first, the
 *          re-runs appended by resolve_strand), list_sfwd / list_srev = reads whose forward- / reverse-strand passes are
 *          score-only; pos_fwd / pos_rev = position of a read in list_fwd / list_rev.
 *   ranges (int32[12], device): {0, n_fwd}, {0, n_rev}, {0, n_sfwd}, {0, n_srev}, {n_fwd, n_fwd_total}, {n_rev, n_rev_total}. */
struct StrandLists {
    int32_t* list_fwd; int32_t* list_rev; int32_t* list_sfwd; int32_t* list_srev;
    int32_t* pos_fwd; int32_t* pos_rev;
    uint8_t* predicted;       /* 0 forward, 1 reverse, 2 unsure */
    int32_t* ranges;
};
struct ClassifyArgs {
    const uint16_t* rows_front; const uint16_t* rows_back;
    const int32_t* lens_front; const int32_t* lens_back;
    long long n;
    int stride;
    const uint32_t* seeds;    /* device, 4096 words: two bits per 8-mer code c (word c >> 4, bits 2 * (c & 15)): occurs in adaptor1 / in adaptor2 */
    int margin;               /* hits by which one strand must lead */
    int scan;                 /* bases of each window that are looked at (the adaptors sit at the windows' start) */
    int vec;                  /* set by the launcher: rows allow 16-byte loads */
    int test_mode;            /* 0; tests: 1 inverts the predictions, 2 calls every read unsure */
    StrandLists L;
};
void launch_classify_strands(const ClassifyArgs& c, cudaStream_t st);

/* Synthetic mockReads-style windows generated on the device (R/mockReads.R:58-92; kernels.cu: mock_windows_kernel). */
struct MockArgs {
    long long n;
    unsigned long long seed, first_index;
    int tol, stride;
    uint16_t* front; uint16_t* back;       /* packed rows [n][stride] */
    int32_t* lens_front; int32_t* lens_back; int32_t* width;
    uint8_t* flipped;                      /* optional */
    int len1, len2;                        /* adaptor lengths (<= 128) */
    int run0_start, run0_end;              /* adaptor1's first N-run (the barcode slot); -1, -1 if none */
    int nbarcodes;                         /* 0: the slot holds one random base repeated (R/mockReads.R:50) */
    const uint8_t* barcodes;               /* device: [nbarcodes][run0_end - run0_start] base codes 0..3 */
    int molecule_len;                      /* adaptor1 + insert + adaptor2 before mutation */
    int max_insert;
    uint32_t sub_thr, indel_thr;           /* 16-bit thresholds: floor(rate * 65536) */
    int wlo;                               /* read width: number of indels = wlo + #{k : word >= wcdf[k]} (Binomial(molecule_len, rate)) */
    uint32_t wcdf[256];                    /* floor(2^32 * P(indels <= wlo + k)) */
    uint32_t qmin;                         /* smallest quality that can occur */
    uint32_t qthr[95];                     /* quality >= k  <=>  word < qthr[k]  (k = 0..93; qthr[94] = 0) */
    char adaptor1[128]; char adaptor2[128];
};
void launch_mock_windows(const MockArgs& m, cudaStream_t st);

/* .scramble_input (R/getAdaptorThresholds.R:68-92) on packed rows: a Fisher-Yates shuffle per window driven by the
 * counter-based stream of (seed, read index, stream_id) (kernels.cu: scramble_rows_fy) -- the same permutation
 * sarlacc_b200/api.py:_scramble_input builds on the host. */
void launch_scramble(const uint16_t* in, uint16_t* out, const int32_t* lens, long long n, int stride,
                     unsigned long long seed, unsigned long long first_index, const unsigned long long* read_index,
                     unsigned long long stream_id, cudaStream_t st);

/* Device packer: raw sequence / quality bytes (as the caller holds them: ASCII or Biostrings codes, any quality
 * encoding) -> the uint16 rows above.  Window i is seq[soff[i] .. soff[i] + lens[i]) and qual[qoff[i] ..); back != 0
 * packs its reverse complement (qualities reversed with it, R/adaptorAlign.R:86-95).  The 256-entry tables are the host
 * packer's (api.cpp: build_pack_tables), so both packers are the same function of the input bytes.  A quality below
 * the encoding's offset (src/reference_align.cpp:215-217) packs index 0 and reports its window: atomicMin(first_bad, i). */
struct PackArgs {
    const uint8_t* seq;
    const uint8_t* qual;
    const long long* soff;
    const long long* qoff;
    const int32_t* lens;
    long long n;
    int stride;
    int back;
    int seq4;                  /* seq holds 4-bit one-hot base codes, two per byte (base k: nibble k & 1 of byte k >> 1), soff counts bases */
    uint16_t* rows;
    long long* first_bad;      /* initialised to LLONG_MAX by the caller; may be null */
    uint8_t base[256];
    uint8_t base_rc[256];
    uint16_t qidx[256];        /* 0xFFFF: below the offset */
};
void launch_pack_rows(const PackArgs& a, cudaStream_t st);

/* Launchers (return the kernel's name for reporting; throw nothing, errors via cudaGetLastError). */
const char* launch_wavefront(const AlignArgs& a, bool trace, bool has_alt, int grid, cudaStream_t st);
const char* launch_generic(const AlignArgs& a, bool trace, int grid, cudaStream_t st);
void launch_traceback(const TraceArgs& t, cudaStream_t st);
void launch_fill_empty(const AlignArgs& a, cudaStream_t st);
size_t wavefront_smem_bytes(const AlignArgs& a);
/* Trace record geometry of the wavefront kernels: bytes of the low word, bytes of the high element (0: none). */
inline int trace_word_bytes(int C) { return C <= 8 ? 4 : 8; }
inline int trace_hi_bytes(int C, bool solo) { return C <= 16 ? 0 : (solo ? 4 : 1); }
inline bool solo_geometry(int G, int C) { return G == 1 && C >= kSoloMinC && C <= kSoloMaxC; }
/* Alignment groups of one full grid of the wavefront kernel on the current device (0: geometry not instantiated).
 * Every group walks its alignments back to back, so a launch over k * groups equal-length alignments has no tail. */
long long wavefront_groups(const AlignArgs& a, bool trace);
int wavefront_block_threads();

}

#endif
