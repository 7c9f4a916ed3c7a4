/* Host side of the C ABI declared in include/sarlacc_b200.h.
 *
 * What lives here (all of it replaces reference host code, cited per function):
 *   - argument validation with the reference's error texts (src/quality_encoding.cpp:5-32,
 *     src/adaptor_align.cpp:23-31,51-53, src/reference_align.cpp:211,215-217);
 *   - cost tables (src/reference_align.cpp:21-52) -- computed on the host with libm log, exactly the
 *     reference's expression, and shipped to the device, so no device transcendental is involved;
 *   - the packer: reads (views or CSR, ASCII or Biostrings byte codes; src/DNA_input.cpp:64-75) ->
 *     2 bytes per base in pinned buffers -> cudaMemcpyAsync;
 *   - chunked, double-buffered execution, sharded over the configured devices by contiguous read
 *     index ranges (the axis .parallelize uses, R/adaptorAlign.R:126-134); no collective is needed;
 *   - the resident (HBM-resident windows) variant.
 * There is deliberately no CPU implementation of the alignment here.
 */
#include "sarlacc_b200.h"
#include "kernels.h"

#include <nvtx3/nvToolsExt.h>     /* header-only; ranges cost nothing unless a profiler is attached */

#include <sys/mman.h>
#include <zlib.h>
#include <sys/types.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <functional>
#include <limits>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace sarlacc;

namespace {

thread_local std::string g_error;
std::atomic<long long> g_launches{0};
std::mutex g_cfg_mutex;
std::vector<int> g_devices;       /* empty = {0} */
int g_host_threads = 0;           /* 0 = auto */

int fail(const std::string& msg) {
    g_error = msg;
    return 1;
}

struct CudaError { std::string msg; };

/* NVTX range over a host-side phase (staging, enqueueing a pass, waiting for results): what a timeline shows next to the
 * kernels of the same phase. */
struct Range {
    explicit Range(const char* name) { nvtxRangePushA(name); }
    ~Range() { nvtxRangePop(); }
    Range(const Range&) = delete;
    Range& operator=(const Range&) = delete;
};

}  // namespace

namespace sarlacc {   /* for the other translation units of the library (umi.cu, threshold.cu) */
int set_error(const std::string& msg) { return fail(msg); }
void count_launches(int n) { g_launches += n; }
void threshold_trim();
}

namespace {

#define CUDA_CHECK(expr)                                                                       \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            throw CudaError{std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #expr}; \
        }                                                                                      \
    } while (0)

/* ---- encoding + cost tables ----------------------------------------------------------------- */

struct Encoding {
    int n = 0;
    char offset = 0;
    std::vector<double> cost;   /* [5][n]: match1, mismatch1, mismatch2, match3, match4 */
};

/* quality_encoding::quality_encoding, src/quality_encoding.cpp:5-32 (messages verbatim). */
const char* check_encoding(const sarlacc_encoding* enc, char* offset) {
    if (!enc || enc->names == nullptr || enc->n <= 0) {
        return "encoding vector must be non-empty and named";
    }
    char last = 0;
    for (int i = 0; i < enc->n; ++i) {
        const char* nm = enc->names[i];
        if (!nm || std::strlen(nm) != 1) {
            return "names of encoding vector must be one character in length";
        }
        const char curval = nm[0];
        if (i > 0) {
            if (curval != last + 1) {   /* int arithmetic on (signed) char, as in the reference */
                return "names of encoding vector should increase consecutively";
            } else if (enc->err[i] > enc->err[i - 1]) {
                return "error probabilities should decrease";
            }
        } else {
            *offset = curval;
        }
        last = curval;
    }
    return nullptr;
}

/* reference_align::create_qualities, src/reference_align.cpp:21-52.  Only the five tables the scoring
 * can reach are kept (src/reference_align.cpp:184-212): match m=1, mismatch m=1, mismatch m=2, match m=3,
 * match m=4.  Expression and evaluation order are the reference's; built with -ffp-contract=off. */
const char* build_encoding(const sarlacc_encoding* enc, Encoding& out) {
    const char* msg = check_encoding(enc, &out.offset);
    if (msg) return msg;
    out.n = enc->n;
    out.cost.assign((size_t)5 * enc->n, 0.0);
    constexpr double n = 4;
    auto match = [&](int i, double epsilon) {
        const double gamma_xy = 1.0 / (i + 1.0);
        const double gamma_xy_1m = 1 - gamma_xy;
        return std::log(gamma_xy * (1 - epsilon) * n + gamma_xy_1m * epsilon * (n / (n - 1))) / M_LN2;
    };
    auto mismatch = [&](int i, double epsilon) {
        const double gamma_xy = 1.0 / (i + 1.0);
        const double gamma_xy_1m = 1 - gamma_xy;
        return std::log(gamma_xy_1m * (1 - epsilon) * n + gamma_xy * epsilon * (n / (n - 1))) / M_LN2;
    };
    for (int j = 0; j < enc->n; ++j) {
        const double e = enc->err[j];
        out.cost[0 * (size_t)enc->n + j] = match(0, e);
        out.cost[1 * (size_t)enc->n + j] = mismatch(0, e);
        out.cost[2 * (size_t)enc->n + j] = mismatch(1, e);
        out.cost[3 * (size_t)enc->n + j] = match(2, e);
        out.cost[4 * (size_t)enc->n + j] = match(3, e);
    }
    return nullptr;
}

/* ---- reference (adaptor / barcode) analysis --------------------------------------------------- */

/* compute_cost's switch, src/reference_align.cpp:184-212.  Returns false for an unrecognized base. */
bool classify_ref(char ref, uint8_t* mask, uint8_t* kind) {
    *mask = 0;
    switch (ref) {
        case 'A': *mask = 1; *kind = COL_ACGT; return true;
        case 'C': *mask = 2; *kind = COL_ACGT; return true;
        case 'G': *mask = 4; *kind = COL_ACGT; return true;
        case 'T': *mask = 8; *kind = COL_ACGT; return true;
        case 'M': case 'R': case 'W': case 'S': case 'Y': case 'K': *kind = COL_TWO; return true;
        case 'V': case 'H': case 'D': case 'B': *kind = COL_THREE; return true;
        case 'N': *kind = COL_N; return true;
    }
    *kind = COL_ACGT;
    return false;
}

struct Plan {
    int L = 0, nref = 1;
    bool local = true;
    double gop = 0, ge = 0;
    const Encoding* enc = nullptr;
    std::vector<double> row0;
    std::vector<uint8_t> refmask, refkind;
    int first_bad_col = -1;       /* first unrecognized reference column over all refs, per ref */
    std::vector<int> bad_col;     /* per reference: first unrecognized column or -1 */
    bool fast = false;
    bool has_alt = false;
    int alt_row = 4;
    int kinds = 1;
    int G = 1, C = 1;         /* lanes per alignment x columns per lane of the wavefront kernels */
    bool solo = false;        /* windows of 48+ rows run one thread per alignment (G = 1, C = L) instead: references of 20-24 bases */
    bool wide_ok = false;     /* score-only runs may use 14-18 columns per lane */
    std::vector<int32_t> sec_starts, sec_ends;
};

bool choose_geometry(Plan& P) {
    const char* force = std::getenv("SARLACC_FORCE_GC");
    P.solo = false;
    if (force) {
        int g = 0, c = 0;
        if (std::sscanf(force, "%d,%d", &g, &c) == 2 && g >= 1 && g <= 32 && (g & (g - 1)) == 0 && c >= 1 && c <= kMaxC && (c <= 12 || !(c & 1)) &&
            g * c >= P.L && g * c - P.L < g) {   /* at most one dummy slot per lane */
            P.G = g;
            P.C = c;
            return true;
        }
    }
    /* 20-24 columns: one thread per alignment, all columns in its registers -- no shuffles, no boundary selects, the
     * row overhead spread over twice the columns of the two-lane split (profiles/r02_history.md) */
    P.solo = P.L >= kSoloMinC && P.L <= kSoloMaxC && std::getenv("SARLACC_NO_SOLO") == nullptr;
    double best = 1e300;
    for (int g = 1; g <= kMaxGroup; g *= 2) {
        const int c = (P.L + g - 1) / g;
        if (c > kMaxC || (c > 12 && (c & 1))) continue;   /* instantiated: 1..12, 14, 16, 18 */
        if (c > 12 && !P.wide_ok) continue;   /* 14-18 columns per lane (3 resident blocks): measured faster only without trace records */
        const double util = (double)P.L / ((double)g * c);
        /* issue slots per DP cell: ~19 for the cell itself + ~86 per row spread over the lane's c columns
         * (profiles/r01_ncu_summary_v2.txt), corrected for the resident warps the register budget allows */
        double est = (19.0 + 86.0 / c) / util;
        if (c > 12) est *= 1.12;                     /* 3 resident blocks per SM */
        else if (c > 9) est *= 1.03;                 /* 4 blocks, but the register cap costs a few spills */
        if (est < best) {
            best = est;
            P.G = g;
            P.C = c;
        }
    }
    return best < 1e300;   /* false: no instantiated geometry covers L (the literal kernel takes it) */
}

void build_plan(Plan& P, const Encoding& enc, const char* const* refs, int nref, int L, bool local, double go, double ge, bool trace = true) {
    P.enc = &enc;
    (void)trace;   /* with two rows per step the 14-18 column geometries win with and without trace records (profiles/) */
    P.wide_ok = std::getenv("SARLACC_NO_WIDE_C") == nullptr;
    P.L = L;
    P.nref = nref;
    P.local = local;
    P.gop = go + ge;     /* src/reference_align.cpp:8 */
    P.ge = ge;
    P.row0.assign((size_t)L + 1, 0.0);
    for (int c = 1; c <= L; ++c) {   /* src/reference_align.cpp:116-118 */
        P.row0[c] = P.row0[c - 1] - (c == 1 ? P.gop : P.ge);
    }
    P.refmask.assign((size_t)nref * L, 0);
    P.refkind.assign((size_t)nref * L, COL_ACGT);
    P.bad_col.assign(nref, -1);
    bool kinds[4] = {false, false, false, false};
    for (int b = 0; b < nref; ++b) {
        for (int c = 0; c < L; ++c) {
            uint8_t m, k;
            if (!classify_ref(refs[b][c], &m, &k)) {
                if (P.bad_col[b] < 0) P.bad_col[b] = c;
            }
            P.refmask[(size_t)b * L + c] = m;
            P.refkind[(size_t)b * L + c] = k;
            kinds[k] = true;
        }
    }
    const int nalt = (int)kinds[COL_TWO] + (int)kinds[COL_THREE] + (int)kinds[COL_N];
    P.has_alt = nalt > 0;
    P.alt_row = kinds[COL_TWO] ? 2 : (kinds[COL_THREE] ? 3 : 4);
    P.kinds = (kinds[COL_ACGT] ? 1 : 0) | (kinds[COL_TWO] ? 2 : 0) | (kinds[COL_THREE] ? 4 : 0) | (kinds[COL_N] ? 8 : 0);
    bool finite = std::isfinite(go) && std::isfinite(ge);
    /* cost tables: -Inf is legitimate (Phred 0: log2(0)) and harmless; NaN / +Inf (malformed error probabilities) would
     * break the max() reformulation, so such encodings take the literal kernel */
    for (double c : enc.cost) {
        if (std::isnan(c) || (std::isinf(c) && c > 0)) finite = false;
    }
    /* The wavefront kernel folds "left/up neighbour already chose a gap" into a max(), which needs
     * gap_open >= gap_ext, i.e. go >= 0 (see kernels.cu).  Anything else takes the literal kernel. */
    P.fast = L >= 1 && L <= kMaxFastL && finite && go >= 0.0 && enc.n <= 256 &&
             std::getenv("SARLACC_FORCE_GENERIC") == nullptr;
    if (P.fast) P.fast = choose_geometry(P);
}

/* ---- read access + packing --------------------------------------------------------------------- */

/* A view of the reads, optionally restricted to the window .get_front_and_back (R/adaptorAlign.R:86-95) would cut:
 * tol > 0, back == false: the first min(tol, width) bases; back == true: the reverse complement of the last
 * min(tol, width) bases (qualities reversed with them) -- produced on the fly by the packer, never materialised. */
struct ReadView {
    const sarlacc_reads* R;
    int tol = 0;
    bool back = false;
    inline int64_t full_seq_len(int64_t i) const { return R->seq ? R->seq_len[i] : R->seq_off[i + 1] - R->seq_off[i]; }
    inline int64_t full_qual_len(int64_t i) const { return R->seq ? R->qual_len[i] : R->qual_off[i + 1] - R->qual_off[i]; }
    inline int64_t clip(int64_t n) const { return (tol > 0 && n > tol) ? tol : n; }
    inline int64_t seq_len(int64_t i) const { return clip(full_seq_len(i)); }
    inline int64_t qual_len(int64_t i) const {
        /* a length mismatch of the whole read must stay visible after clipping */
        const int64_t s = full_seq_len(i), q = full_qual_len(i);
        return (s == q) ? clip(q) : (clip(s) == clip(q) ? clip(q) + 1 : clip(q));
    }
    inline const uint8_t* seq(int64_t i) const {
        const uint8_t* p = R->seq ? R->seq[i] : R->seq_pool + R->seq_off[i];
        return back ? p + (full_seq_len(i) - seq_len(i)) : p;
    }
    inline const uint8_t* qual(int64_t i) const {
        const uint8_t* p = R->seq ? R->qual[i] : R->qual_pool + R->qual_off[i];
        return back ? p + (full_qual_len(i) - clip(full_qual_len(i))) : p;
    }
};

struct PackTables {
    uint8_t base[256];
    uint8_t base_rc[256];   /* one-hot code of the complement (A<->T, C<->G); 0 for everything else */
    /* quality byte -> index, or 0xFFFF when below the offset (signed char comparison, :215) */
    uint16_t qidx[256];
    /* the same mapping in closed form for the vector packer */
    uint8_t code[4];
    int8_t offset;
    uint8_t clamp;
    bool simd_ok;
};

void build_pack_tables(PackTables& T, int seq_encoding, const Encoding& enc) {
    std::memset(T.base, 0, sizeof(T.base));
    if (seq_encoding == SARLACC_SEQ_BIOSTRINGS) {
        T.base[1] = 1; T.base[2] = 2; T.base[4] = 4; T.base[8] = 8;   /* DNAdecode: A C G T */
    } else {
        T.base[(unsigned char)'A'] = 1; T.base[(unsigned char)'C'] = 2;
        T.base[(unsigned char)'G'] = 4; T.base[(unsigned char)'T'] = 8;
    }
    if (seq_encoding == SARLACC_SEQ_BIOSTRINGS) {
        T.code[0] = 1; T.code[1] = 2; T.code[2] = 4; T.code[3] = 8;
    } else {
        T.code[0] = 'A'; T.code[1] = 'C'; T.code[2] = 'G'; T.code[3] = 'T';
    }
    T.offset = (int8_t)enc.offset;
    T.clamp = (uint8_t)std::min(enc.n - 1, 255);
    T.simd_ok = enc.n >= 1;
    static const uint8_t comp[16] = {0, 8, 4, 0, 2, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0};
    for (int b = 0; b < 256; ++b) T.base_rc[b] = comp[T.base[b] & 15];
    for (int b = 0; b < 256; ++b) {
        const char q = (char)(unsigned char)b;
        if (q < enc.offset) {
            T.qidx[b] = 0xFFFF;
        } else {
            size_t loc = (size_t)(q - enc.offset);
            if (loc >= (size_t)enc.n) loc = (size_t)enc.n - 1;   /* :219-221 */
            T.qidx[b] = (uint16_t)loc;
        }
    }
}

enum ErrKind { ERR_NONE = 0, ERR_LEN, ERR_QUAL, ERR_REF };

struct FirstError {
    int64_t at = std::numeric_limits<int64_t>::max();
    ErrKind kind = ERR_NONE;
    void offer(int64_t i, ErrKind k) {
        if (i < at) { at = i; kind = k; }
    }
    void merge(const FirstError& o) {
        if (o.kind != ERR_NONE) offer(o.at, o.kind);
    }
};

const char* err_text(ErrKind k) {
    switch (k) {
        case ERR_LEN: return "sequence and quality strings should have the same length";
        case ERR_QUAL: return "quality cannot be lower than smallest encoded value";
        case ERR_REF: return "unrecognized base in reference sequence";
        default: return "";
    }
}

int host_threads_for(int ndev) {
    int t = g_host_threads;
    if (t <= 0) {
        t = (int)std::thread::hardware_concurrency();
        if (t <= 0) t = 4;
        /* one process per GPU (torchrun): the ranks of a node share its cores */
        const char* lws = std::getenv("LOCAL_WORLD_SIZE");
        const int ranks = lws ? std::max(1, std::atoi(lws)) : 1;
        t = std::max(1, t / (std::max(1, ndev) * ranks));
        t = std::min(t, 32);
    }
    return t;
}

/* Persistent worker pool for the packer (thread creation per chunk showed up in the end-to-end profile). */
class WorkerPool {
public:
    static WorkerPool& instance() {
        static WorkerPool pool;
        return pool;
    }
    /* Runs fn(k) for k in [0, ntasks) on the pool (the caller takes part) and returns when all are done. */
    void run(int ntasks, const std::function<void(int)>& fn) {
        if (ntasks <= 1) {
            if (ntasks == 1) fn(0);
            return;
        }
        auto job = std::make_shared<Job>();
        job->fn = &fn;
        job->total = ntasks;
        {
            std::lock_guard<std::mutex> lock(m_);
            grow(ntasks - 1);
            queue_.push_back(job);
        }
        cv_.notify_all();
        work_on(*job);
        std::unique_lock<std::mutex> lock(job->m);
        job->cv.wait(lock, [&] { return job->done == job->total; });
    }

private:
    struct Job {
        const std::function<void(int)>* fn = nullptr;
        int total = 0;
        std::atomic<int> next{0};
        int done = 0;
        std::mutex m;
        std::condition_variable cv;
    };
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<std::shared_ptr<Job> > queue_;
    std::vector<std::thread> threads_;
    bool stop_ = false;

    WorkerPool() {}
    ~WorkerPool() {
        {
            std::lock_guard<std::mutex> lock(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : threads_) t.join();
    }
    void grow(int want) {   /* m_ held */
        want = std::min(want, 64);
        while ((int)threads_.size() < want) threads_.emplace_back([this] { loop(); });
    }
    static void work_on(Job& job) {
        for (;;) {
            const int k = job.next.fetch_add(1);
            if (k >= job.total) break;
            (*job.fn)(k);
            std::lock_guard<std::mutex> lock(job.m);
            if (++job.done == job.total) job.cv.notify_all();
        }
    }
    void loop() {
        for (;;) {
            std::shared_ptr<Job> job;
            {
                std::unique_lock<std::mutex> lock(m_);
                cv_.wait(lock, [&] { return stop_ || !queue_.empty(); });
                if (stop_) return;
                job = queue_.front();
                if (job->next.load() >= job->total) {
                    queue_.pop_front();
                    continue;
                }
            }
            work_on(*job);
        }
    }
};

template <class F>
void parallel_for(int64_t lo, int64_t hi, int nthreads, F body) {
    const int64_t n = hi - lo;
    if (n <= 0) return;
    if (nthreads > n / 2048 + 1) nthreads = (int)(n / 2048 + 1);
    if (nthreads <= 1) {
        body(lo, hi, 0);
        return;
    }
    const std::function<void(int)> task = [&](int t) {
        const int64_t a = lo + n * t / nthreads, b = lo + n * (t + 1) / nthreads;
        body(a, b, t);
    };
    WorkerPool::instance().run(nthreads, task);
}

/* Pass 1 over [lo,hi): lengths, length-mismatch errors, max length. */
void scan_lengths(const ReadView& V, int64_t lo, int64_t hi, int32_t* lens, int nthreads, FirstError& err, int& maxlen) {
    std::vector<FirstError> errs(nthreads + 1);
    std::vector<int> mx(nthreads + 1, 0);
    parallel_for(lo, hi, nthreads, [&](int64_t a, int64_t b, int t) {
        for (int64_t i = a; i < b; ++i) {
            const int64_t sl = V.seq_len(i);
            if (sl != V.qual_len(i)) {
                errs[t].offer(i, ERR_LEN);
                lens[i - lo] = 0;
                continue;
            }
            lens[i - lo] = (int32_t)sl;
            if (sl > mx[t]) mx[t] = (int)sl;
        }
    });
    maxlen = 0;
    for (int t = 0; t <= nthreads; ++t) {
        err.merge(errs[t]);
        maxlen = std::max(maxlen, mx[t]);
    }
}

/* AVX2 packer (chosen at run time): 32 bases per iteration.  one-hot base = OR of four byte compares against the
 * encoding's A/C/G/T codes; quality index = min(q - offset, |enc| - 1) with the reference's signed `q < offset` test
 * (src/reference_align.cpp:215-221); the two byte vectors are interleaved into the uint16 rows.  The tail re-does the
 * last 32 bases of the window (overlapping stores of identical values) instead of a scalar loop. */
#if defined(__x86_64__) && defined(__GNUC__)
#define SARLACC_HAVE_AVX2_PACK 1
#include <immintrin.h>

struct PackConsts {
    uint8_t code[4];      /* input bytes of A, C, G, T */
    int8_t offset;        /* smallest encoded quality (signed char, like the reference) */
    uint8_t clamp;        /* min(|enc| - 1, 255) */
};

__attribute__((target("avx2"))) static inline void pack32_avx2(__m256i x, __m256i qv, const PackConsts& K, bool rc, uint16_t* out, __m256i& bad_acc) {
    const __m256i b0 = _mm256_and_si256(_mm256_cmpeq_epi8(x, _mm256_set1_epi8((char)K.code[0])), _mm256_set1_epi8(rc ? 8 : 1));
    const __m256i b1 = _mm256_and_si256(_mm256_cmpeq_epi8(x, _mm256_set1_epi8((char)K.code[1])), _mm256_set1_epi8(rc ? 4 : 2));
    const __m256i b2 = _mm256_and_si256(_mm256_cmpeq_epi8(x, _mm256_set1_epi8((char)K.code[2])), _mm256_set1_epi8(rc ? 2 : 4));
    const __m256i b3 = _mm256_and_si256(_mm256_cmpeq_epi8(x, _mm256_set1_epi8((char)K.code[3])), _mm256_set1_epi8(rc ? 1 : 8));
    const __m256i base = _mm256_or_si256(_mm256_or_si256(b0, b1), _mm256_or_si256(b2, b3));
    const __m256i offv = _mm256_set1_epi8((char)K.offset);
    const __m256i bad = _mm256_cmpgt_epi8(offv, qv);                                /* signed q < offset */
    bad_acc = _mm256_or_si256(bad_acc, bad);
    __m256i qi = _mm256_min_epu8(_mm256_sub_epi8(qv, offv), _mm256_set1_epi8((char)K.clamp));
    qi = _mm256_andnot_si256(bad, qi);                                              /* failing reads pack index 0 */
    const __m256i lo = _mm256_unpacklo_epi8(qi, base);                              /* elements 0-7 | 16-23 */
    const __m256i hi = _mm256_unpackhi_epi8(qi, base);                              /* elements 8-15 | 24-31 */
    _mm256_storeu_si256(reinterpret_cast<__m256i*>(out), _mm256_permute2x128_si256(lo, hi, 0x20));
    _mm256_storeu_si256(reinterpret_cast<__m256i*>(out + 16), _mm256_permute2x128_si256(lo, hi, 0x31));
}

__attribute__((target("avx2"))) static inline __m256i reverse32_avx2(__m256i v) {
    const __m256i idx = _mm256_setr_epi8(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
    return _mm256_permute2x128_si256(_mm256_shuffle_epi8(v, idx), _mm256_shuffle_epi8(v, idx), 0x01);
}

/* One window of len >= 32 bases; returns true if some quality was below the offset. */
__attribute__((target("avx2"))) static bool pack_window_avx2(const uint8_t* s, const uint8_t* q, int len, bool back, const PackConsts& K, uint16_t* out) {
    __m256i bad = _mm256_setzero_si256();
    for (int r = 0;; r += 32) {
        if (r + 32 > len) r = len - 32;      /* last block: overlap */
        if (!back) {
            pack32_avx2(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + r)),
                        _mm256_loadu_si256(reinterpret_cast<const __m256i*>(q + r)), K, false, out + r, bad);
        } else {
            /* out[r .. r+31] come from positions len-1-r down to len-32-r */
            pack32_avx2(reverse32_avx2(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + len - 32 - r))),
                        reverse32_avx2(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(q + len - 32 - r))), K, true, out + r, bad);
        }
        if (r + 32 >= len) break;
    }
    return !_mm256_testz_si256(bad, bad);
}

static bool have_avx2() {
    static const bool ok = __builtin_cpu_supports("avx2") && std::getenv("SARLACC_NO_AVX2") == nullptr;
    return ok;
}
#endif

/* Pass 2: rows[(i-lo)*stride + r] = qidx | base << 8.  Quality errors only matter when a cost would have been
 * computed for that read, i.e. L > 0 (src/reference_align.cpp:184-225 is only reached from align_column). */
void pack_rows(const ReadView& V, int64_t lo, int64_t hi, const PackTables& T, const int32_t* lens, int stride,
        uint16_t* rows, int nthreads, bool check_qual, FirstError& err, bool force_scalar = false)
{
    std::vector<FirstError> errs(nthreads + 1);
#ifdef SARLACC_HAVE_AVX2_PACK
    PackConsts K;
    bool simd = have_avx2() && T.simd_ok && !force_scalar;
    if (simd) {
        std::memcpy(K.code, T.code, 4);
        K.offset = T.offset;
        K.clamp = T.clamp;
    }
#endif
    parallel_for(lo, hi, nthreads, [&](int64_t a, int64_t b, int t) {
        for (int64_t i = a; i < b; ++i) {
            const int len = lens[i - lo];
            uint16_t* out = rows + (size_t)(i - lo) * stride;
            const uint8_t* s = V.seq(i);
            const uint8_t* q = V.qual(i);
#ifdef SARLACC_HAVE_AVX2_PACK
            if (simd && len >= 32) {
                if (pack_window_avx2(s, q, len, V.back, K, out) && check_qual) errs[t].offer(i, ERR_QUAL);
                continue;
            }
#endif
            unsigned bad = 0;
            if (!V.back) {
                for (int r = 0; r < len; ++r) {
                    const unsigned qi = T.qidx[q[r]];
                    bad |= qi;   /* 0xFFFF marks a quality below the offset: the call fails; pack index 0 so no kernel reads out of its tables */
                    out[r] = (uint16_t)(((qi & 0xFF00u) ? 0u : qi) | ((unsigned)T.base[s[r]] << 8));
                }
            } else {
                for (int r = 0; r < len; ++r) {
                    const unsigned qi = T.qidx[q[len - 1 - r]];
                    bad |= qi;
                    out[r] = (uint16_t)(((qi & 0xFF00u) ? 0u : qi) | ((unsigned)T.base_rc[s[len - 1 - r]] << 8));
                }
            }
            if (check_qual && (bad & 0xFF00u)) errs[t].offer(i, ERR_QUAL);
        }
    });
    for (int t = 0; t <= nthreads; ++t) err.merge(errs[t]);
}

/* ---- device-side resources ------------------------------------------------------------------- */

/* Freed device buffers are kept for reuse (per device, best fit) instead of going back to the driver: building and
 * dropping resident objects per batch otherwise spends more time in cudaMalloc / cudaFree of multi-GB trace scratch than
 * in the kernels.  A buffer enters the pool only after its device has been synchronised -- the guarantee cudaFree gave.
 * SARLACC_POOL_MB bounds what is kept (default 24576; 0 disables the pool); sarlacc_trim_device_memory() empties it. */
class DevPool {
public:
    static DevPool& instance() {
        static DevPool* pool = new DevPool();     /* never destroyed: no CUDA calls during process teardown */
        return *pool;
    }
    void* take(size_t bytes, int dev, size_t* cap) {
        std::lock_guard<std::mutex> lock(m_);
        int best = -1;
        for (size_t i = 0; i < free_.size(); ++i) {
            const Block& b = free_[i];
            if (b.dev != dev || b.cap < bytes || b.cap > 2 * bytes + (1u << 20)) continue;
            if (best < 0 || b.cap < free_[(size_t)best].cap) best = (int)i;
        }
        if (best < 0) return nullptr;
        void* p = free_[(size_t)best].p;
        *cap = free_[(size_t)best].cap;
        cached_ -= *cap;
        free_.erase(free_.begin() + best);
        return p;
    }
    /* returns false if the block was not kept (the caller frees it) */
    bool give(void* p, size_t cap, int dev) {
        std::lock_guard<std::mutex> lock(m_);
        if (limit_ == 0 || cap > limit_) return false;
        while (cached_ + cap > limit_ && !free_.empty()) {      /* make room: drop the oldest blocks */
            cudaFree(free_.front().p);
            cached_ -= free_.front().cap;
            free_.erase(free_.begin());
        }
        free_.push_back(Block{p, cap, dev});
        cached_ += cap;
        return true;
    }
    void trim() {
        std::lock_guard<std::mutex> lock(m_);
        for (auto& b : free_) cudaFree(b.p);
        free_.clear();
        cached_ = 0;
    }

private:
    struct Block { void* p; size_t cap; int dev; };
    DevPool() {
        const char* e = std::getenv("SARLACC_POOL_MB");
        long mb = 24576;
        if (e) mb = std::atol(e);
        limit_ = mb <= 0 ? 0 : (size_t)mb << 20;
    }
    std::mutex m_;
    std::vector<Block> free_;
    size_t cached_ = 0, limit_ = 0;
};

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int dev = 0;
    void reserve(size_t bytes) {
        if (bytes <= cap) return;
        release();
        CUDA_CHECK(cudaGetDevice(&dev));
        const size_t want = bytes + bytes / 8 + 256;
        p = DevPool::instance().take(want, dev, &cap);
        if (p) return;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {          /* the pool may be what fills the device: empty it and try once more */
            cudaGetLastError();
            DevPool::instance().trim();
            p = nullptr;
            CUDA_CHECK(cudaMalloc(&p, want));
        }
        cap = want;
    }
    void release() {
        if (p) {
            int cur = 0;
            cudaGetDevice(&cur);
            if (cur != dev) cudaSetDevice(dev);
            cudaDeviceSynchronize();     /* nothing in flight may still touch the block when someone else takes it */
            if (!DevPool::instance().give(p, cap, dev)) cudaFree(p);
            if (cur != dev) cudaSetDevice(cur);
        }
        p = nullptr;
        cap = 0;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    void reserve(size_t bytes) {
        if (bytes <= cap) return;
        if (p) CUDA_CHECK(cudaFreeHost(p));
        p = nullptr;
        cap = 0;
        const size_t want = bytes + bytes / 8 + 256;
        CUDA_CHECK(cudaMallocHost(&p, want));
        cap = want;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

/* Device copy of a Plan's small tables. */
struct DevPlan {
    DevBuf buf;
    const double* row0 = nullptr;
    const double* cost = nullptr;
    const uint8_t* refmask = nullptr;
    const uint8_t* refkind = nullptr;
    const int32_t* sec_starts = nullptr;
    const int32_t* sec_ends = nullptr;

    void upload(const Plan& P, cudaStream_t st) {
        auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };
        const size_t o_row0 = 0;
        const size_t o_cost = o_row0 + al(sizeof(double) * P.row0.size());
        const size_t o_mask = o_cost + al(sizeof(double) * P.enc->cost.size());
        const size_t o_kind = o_mask + al(P.refmask.size() + 1);
        const size_t o_ss = o_kind + al(P.refkind.size() + 1);
        const size_t o_se = o_ss + al(sizeof(int32_t) * (P.sec_starts.size() + 1));
        const size_t total = o_se + al(sizeof(int32_t) * (P.sec_ends.size() + 1));
        std::vector<uint8_t> h(total, 0);
        std::memcpy(h.data() + o_row0, P.row0.data(), sizeof(double) * P.row0.size());
        std::memcpy(h.data() + o_cost, P.enc->cost.data(), sizeof(double) * P.enc->cost.size());
        if (!P.refmask.empty()) std::memcpy(h.data() + o_mask, P.refmask.data(), P.refmask.size());
        if (!P.refkind.empty()) std::memcpy(h.data() + o_kind, P.refkind.data(), P.refkind.size());
        if (!P.sec_starts.empty()) {
            std::memcpy(h.data() + o_ss, P.sec_starts.data(), sizeof(int32_t) * P.sec_starts.size());
            std::memcpy(h.data() + o_se, P.sec_ends.data(), sizeof(int32_t) * P.sec_ends.size());
        }
        buf.reserve(total);
        CUDA_CHECK(cudaMemcpyAsync(buf.p, h.data(), total, cudaMemcpyHostToDevice, st));
        CUDA_CHECK(cudaStreamSynchronize(st));   /* h is a stack-owned staging vector */
        const uint8_t* base = buf.as<uint8_t>();
        row0 = reinterpret_cast<const double*>(base + o_row0);
        cost = reinterpret_cast<const double*>(base + o_cost);
        refmask = base + o_mask;
        refkind = base + o_kind;
        sec_starts = reinterpret_cast<const int32_t*>(base + o_ss);
        sec_ends = reinterpret_cast<const int32_t*>(base + o_se);
    }
};

struct Outputs {   /* device pointers; any may be null */
    double* score = nullptr;       /* [nref][n] */
    int32_t* best_id = nullptr;
    double* best = nullptr;
    double* next_best = nullptr;
    int32_t* start = nullptr;
    int32_t* end = nullptr;
    int32_t* sec_start = nullptr;  /* [nsec][n] */
    int32_t* sec_width = nullptr;
    uint8_t* ops = nullptr;
    int32_t* nops = nullptr;
    long long ops_stride = 0;
};

/* Scratch for one in-flight run on one stream. */
struct Scratch {
    DevBuf flags, map, endrow, gS, gE, gC, next;
    void release() { flags.release(); map.release(); endrow.release(); gS.release(); gE.release(); gC.release(); next.release(); }
};

size_t scratch_budget_bytes() {
    const char* e = std::getenv("SARLACC_SCRATCH_MB");
    size_t mb = 4096;
    if (e) {
        const long v = std::atol(e);
        if (v >= 16) mb = (size_t)v;
    }
    return mb << 20;
}

/* Two rows per lane step (wf_forward2) pay off on long windows; on barcode-length reads (a couple of dozen rows per
 * alignment, lanes of a warp at different phases of different alignments) nearly every step is the masked one and the
 * single-row kernel is ~2x faster (1 M x 96 barcodes: 0.11 s vs 0.20 s).  SARLACC_PAIR=0/1 forces either (A/B tests). */
int pair_rows_default(int maxlen = 1 << 30) {
    const char* e = std::getenv("SARLACC_PAIR");
    if (e) return std::atoi(e);
    return maxlen >= 48 ? 1 : 0;
}

/* The geometry a run over windows of at most `maxlen` rows uses: the solo kernel is a row-pair kernel whose lanes are
 * independent alignments, so it wants windows long enough to stay in step (the same bound as pair_rows_default). */
struct Geometry { int G, C; bool solo; int pair; };
Geometry geometry_for(const Plan& P, int maxlen = 1 << 30, bool by_length = false) {
    /* by_length: the launch walks its reads in order of length, which keeps a warp's alignments in step however short
     * they are, so barcode-length reads can take the solo kernel too.  Measured on configs[3] (1 M x 96 barcodes,
     * profiles/r02_history.md): ordering alone 0.106 -> 0.074 s with the two-lane one-row kernel, 0.066 s with the solo
     * kernel (SARLACC_SOLO_SHORT=0 keeps the two-lane kernel, for comparison). */
    static const bool solo_short = [] {
        const char* e = std::getenv("SARLACC_SOLO_SHORT");
        return e ? std::atoi(e) != 0 : true;
    }();
    if (P.solo && (pair_rows_default(maxlen) ? P.nref == 1 : (by_length && solo_short))) return Geometry{1, P.L, true, 1};
    return Geometry{P.G, P.C, false, pair_rows_default(maxlen)};
}

/* How many alignments of at most `maxlen` rows one sub-launch may cover under the scratch budget. */
long long sub_chunk(const Plan& P, int maxlen, bool trace, long long n) {
    size_t per = 0;
    if (trace) {
        if (P.fast) {
            const Geometry g = geometry_for(P, maxlen);
            const size_t wb = (size_t)trace_word_bytes(g.C) + (size_t)trace_hi_bytes(g.C, g.solo);
            per += g.solo ? (size_t)(maxlen + 1) * wb : (size_t)(maxlen + kSkew * g.G) * g.G * wb;
        } else {
            per += (size_t)std::max(1, maxlen) * P.L;
        }
        per += sizeof(int32_t) * ((size_t)P.L + 2);
    }
    if (per == 0) return n;
    long long cn = (long long)(scratch_budget_bytes() / per);
    if (cn < 1) cn = 1;
    return std::min(cn, n);
}

const char* g_last_kernel = "";

/* Alignment groups in one full grid of the plan's forward kernel (0 for the literal kernel). */
long long plan_groups(const Plan& P, bool trace) {
    if (!P.fast) return 0;
    const Geometry g = geometry_for(P);
    AlignArgs A;
    std::memset(&A, 0, sizeof(A));
    A.L = P.L;
    A.nref = P.nref;
    A.enc_n = P.enc->n;
    A.G = g.G;
    A.C = g.C;
    A.solo = g.solo ? 1 : 0;
    A.pair_rows = g.pair;
    return wavefront_groups(A, trace);
}

/* The same for the geometry a launch over windows of at most `maxlen` rows takes. */
long long launch_groups(const Plan& P, int maxlen, bool by_length, bool trace) {
    if (!P.fast) return 0;
    const Geometry g = geometry_for(P, maxlen, by_length);
    AlignArgs A;
    std::memset(&A, 0, sizeof(A));
    A.L = P.L;
    A.nref = P.nref;
    A.enc_n = P.enc->n;
    A.G = g.G;
    A.C = g.C;
    A.solo = g.solo ? 1 : 0;
    A.pair_rows = g.pair;
    return wavefront_groups(A, trace);
}

/* Every group walks its alignments back to back, so a launch over a whole number of grid-fulls of (equal-length)
 * alignments has no tail round: 131 072 windows on 14 208 groups would run 10 rounds for 9.2 rounds of work. */
long long whole_rounds(long long cn, long long groups) {
    if (groups <= 0 || cn < groups) return cn;
    return cn / groups * groups;
}

/* Reads per pipeline chunk near `target` that is a whole number of rounds for both plans if possible, else for the first. */
long long chunk_for(long long target, long long g1, long long g2) {
    if (g1 <= 0 && g2 <= 0) return target;
    if (g1 <= 0) std::swap(g1, g2);
    if (g2 > 0) {
        long long a = g1, b = g2;
        while (b) { const long long t = a % b; a = b; b = t; }
        const long long l = g1 / a * g2;
        if (l <= 2 * target) return std::max<long long>(1, (target + l / 2) / l) * l;
    }
    return std::max<long long>(1, (target + g1 / 2) / g1) * g1;
}

/* Optional CUDA-event bracket around the forward kernel launches of one run (roofline accounting). */
struct FwdTimer {
    std::vector<cudaEvent_t> ev;   /* pairs */
    size_t used = 0;
    void begin(cudaStream_t st) {
        if (used + 2 > ev.size()) {
            cudaEvent_t a, b;
            CUDA_CHECK(cudaEventCreate(&a));
            CUDA_CHECK(cudaEventCreate(&b));
            ev.push_back(a);
            ev.push_back(b);
        }
        CUDA_CHECK(cudaEventRecord(ev[used], st));
    }
    void end(cudaStream_t st) {
        CUDA_CHECK(cudaEventRecord(ev[used + 1], st));
        used += 2;
    }
    double total_ms() {
        double t = 0;
        for (size_t i = 0; i + 1 < used; i += 2) {
            CUDA_CHECK(cudaEventSynchronize(ev[i + 1]));
            float ms = 0;
            CUDA_CHECK(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
            t += ms;
        }
        return t;
    }
    void release() {
        for (auto e : ev) cudaEventDestroy(e);
        ev.clear();
        used = 0;
    }
};

bool dynamic_distribution() {
    static const bool on = [] {
        const char* e = std::getenv("SARLACC_DYNAMIC");
        return e ? std::atoi(e) != 0 : true;
    }();
    return on;
}

/* One forward launch over device-resident packed windows [0, m) (+ the scores of empty windows), with traceback records
 * into S when `trace`.  Returns what a traceback of these records needs. */
struct FwdRec {
    AlignArgs A;
    Geometry geo;
    int layout = 0, wordbytes = 0, hi_bytes = 0;
    const char* name = "";
};

/* Which reads a forward launch covers and in what order: a list of read indices (null: all, in order), device-side
 * bounds of the part of the list to take (null: all of it), the device counter of the dynamic distribution (null: the
 * launch's own), a cap on the grid (0: a full grid). */
struct FwdOrder {
    const int32_t* index = nullptr;
    const int32_t* range = nullptr;
    unsigned long long* next = nullptr;
    int grid = 0;
    bool fill_empty = true;      /* false: another launch over the same reads already wrote the empty reads' scores */
};

FwdRec forward_once(const Plan& P, const DevPlan& D, Scratch& S, cudaStream_t st,
        const uint16_t* d_rows, const int32_t* d_lens, long long m, int stride, int maxlen,
        bool trace, const Outputs& out, int sms, FwdTimer* timer = nullptr, const int32_t* d_by_length = nullptr,
        const FwdOrder* order = nullptr)
{
    Range nvtx(trace ? "sarlacc: forward pass (trace)" : "sarlacc: forward pass (score)");
    FwdRec R;
    AlignArgs& A = R.A;
    std::memset(&A, 0, sizeof(A));
    A.rows = d_rows;
    A.lens = d_lens;
    A.n = m;
    A.stride = stride;
    A.L = P.L;
    A.nref = P.nref;
    A.refmask = D.refmask;
    A.refkind = D.refkind;
    A.local = P.local ? 1 : 0;
    A.gop = P.gop;
    A.ge = P.ge;
    A.row0 = D.row0;
    A.cost = D.cost;
    A.enc_n = P.enc->n;
    A.kinds = P.kinds;
    A.one = 1u;
    const Geometry geo = geometry_for(P, maxlen, d_by_length != nullptr);
    R.geo = geo;
    A.index = order ? order->index : d_by_length;
    A.G = geo.G;
    A.C = geo.C;
    A.solo = geo.solo ? 1 : 0;
    A.pair_rows = geo.pair;
    A.score = out.score;
    A.best_id = out.best_id;
    A.best = out.best;
    A.next_best = out.next_best;
    if (trace) {
        if (P.fast) {
            const int wb = trace_word_bytes(geo.C), hb = trace_hi_bytes(geo.C, geo.solo);
            /* record words of this launch: wavefront [alignment][row slot][lane]; solo [32 alignments][row][lane] */
            A.fstride = geo.solo ? (long long)(maxlen + 1) : (long long)(maxlen + kSkew * geo.G) * geo.G;
            const size_t words = geo.solo ? (size_t)((m + 31) / 32) * 32 * (size_t)A.fstride : (size_t)A.fstride * (size_t)m;
            const size_t lo_bytes = (words * wb + 255) & ~(size_t)255;
            S.flags.reserve(lo_bytes + words * hb);
            S.endrow.reserve(sizeof(int32_t) * (size_t)m);
            A.endrow = std::getenv("SARLACC_NO_ENDROW") ? nullptr : S.endrow.as<int32_t>();   /* A/B switch for profiles */
            A.flags_hi = hb ? S.flags.as<uint8_t>() + lo_bytes : nullptr;
            R.layout = geo.solo ? 2 : 0;
            R.wordbytes = wb;
            R.hi_bytes = hb;
        } else {
            A.fstride = (long long)std::max(1, maxlen) * P.L;
            S.flags.reserve((size_t)A.fstride * m);
            R.layout = 1;
            R.wordbytes = 1;
        }
        A.flags = S.flags.p;
    }
    if (order && order->next) {
        A.next = order->next;          /* zeroed by the caller */
        A.range = order->range;
    } else if (P.fast && geo.pair && P.nref == 1 && dynamic_distribution()) {
        /* row-pair kernels take their alignments from a device counter: no tail round whatever the launch's length */
        S.next.reserve(sizeof(unsigned long long));
        CUDA_CHECK(cudaMemsetAsync(S.next.p, 0, sizeof(unsigned long long), st));
        A.next = S.next.as<unsigned long long>();
    }
    if (timer) timer->begin(st);
    if (P.fast) {
        R.name = launch_wavefront(A, trace, P.has_alt, order ? order->grid : 0, st);
    } else {
        long long threads = std::min<long long>(m, (long long)sms * 1024);
        threads = (threads + 127) / 128 * 128;
        A.gthreads = threads;
        S.gS.reserve(sizeof(double) * (size_t)(maxlen + 1) * threads);
        S.gE.reserve(sizeof(double) * (size_t)(maxlen + 1) * threads);
        S.gC.reserve((size_t)(maxlen + 1) * threads);
        A.gS = S.gS.as<double>();
        A.gE = S.gE.as<double>();
        A.gChoice = S.gC.as<uint8_t>();
        R.name = launch_generic(A, trace, 0, st);
    }
    if (timer) timer->end(st);
    if (!order || order->fill_empty) {
        launch_fill_empty(A, st);
        g_launches += 1;
    }
    g_launches += 1;
    return R;
}

/* The traceback arguments that read the records of one forward launch (outputs still to be filled in). */
TraceArgs trace_args_for(const Plan& P, const DevPlan& D, const FwdRec& R) {
    TraceArgs T;
    std::memset(&T, 0, sizeof(T));
    T.lens = R.A.lens;
    T.n = R.A.n;
    T.L = P.L;
    T.layout = R.layout;
    T.G = R.geo.G;
    T.C = R.geo.C;
    T.wordbytes = R.wordbytes;
    T.hi_bytes = R.hi_bytes;
    T.flags = R.A.flags;
    T.flags_hi = R.A.flags_hi;
    T.fstride = R.A.fstride;
    T.endrow = R.A.endrow;
    T.nsec = (int)P.sec_starts.size();
    T.sec_starts = D.sec_starts;
    T.sec_ends = D.sec_ends;
    return T;
}

/* Enqueue forward (+ traceback) for device-resident packed reads [0,n).  The caller hands over a range one launch can
 * take (sub_chunk() under the scratch budget): the [nref][n] / [nsec][n] output matrices have row pitch n. */
const char* run_device(const Plan& P, const DevPlan& D, Scratch& S, cudaStream_t st,
        const uint16_t* d_rows, const int32_t* d_lens, long long n, int stride, int maxlen,
        bool trace, const Outputs& out, int sms, FwdTimer* timer = nullptr,
        cudaStream_t tb_stream = nullptr, cudaEvent_t fwd_done = nullptr, cudaEvent_t tb_done = nullptr,
        const int32_t* d_by_length = nullptr)
{
    /* With tb_stream set, the traceback of this range runs on that stream after fwd_done (recorded on `st`), and
     * tb_done is recorded behind it: the memory-bound traceback then overlaps the next range's ALU-bound forward pass
     * (it fits beside the forward kernel's resident blocks: 32 registers per thread).  The caller owns the waits that
     * protect the scratch buffers. */
    if (n == 0) return "";
    if (sub_chunk(P, maxlen, trace, n) < n) throw CudaError{"internal error: run_device was handed more alignments than its scratch budget covers"};
    const FwdRec R = forward_once(P, D, S, st, d_rows, d_lens, n, stride, maxlen, trace, out, sms, timer, trace ? nullptr : d_by_length);
    if (trace) {
        TraceArgs T = trace_args_for(P, D, R);
        S.map.reserve(sizeof(int32_t) * ((size_t)P.L + 1) * (size_t)n);
        T.map = S.map.as<int32_t>();
        T.start = out.start;
        T.end = out.end;
        T.sec_start = out.sec_start;
        T.sec_width = out.sec_width;
        T.ops = out.ops;
        T.nops = out.nops;
        T.ops_stride = out.ops_stride;
        if (tb_stream) {
            CUDA_CHECK(cudaEventRecord(fwd_done, st));
            CUDA_CHECK(cudaStreamWaitEvent(tb_stream, fwd_done, 0));
            launch_traceback(T, tb_stream);
            CUDA_CHECK(cudaEventRecord(tb_done, tb_stream));
        } else {
            launch_traceback(T, st);
        }
        g_launches += 1;
    }
    CUDA_CHECK(cudaGetLastError());
    g_last_kernel = R.name;
    return R.name;
}

/* ---- both adaptors on both window sets of device-resident reads ---------------------------------------------------
 * .align_AA_internal (R/adaptorAlign.R:178-199) without the work R throws away: the four forward passes -- (adaptor1,
 * front), (adaptor2, back), (adaptor1, back), (adaptor2, front) -- keep their traceback records, .resolve_strand
 * (:112-122) runs on the four score vectors, and then ONE traceback per adaptor walks, for every read, the records of
 * the strand that was kept (cur.starts[rev,] <- cur.rc.starts[rev,], :195-196), writing the final columns directly
 * (adaptor2's start/end flipped into read coordinates when widths are given, :66-71).  Half the traceback work of four
 * separate adaptor_align calls, and no per-strand result sets to select from afterwards. */
/* Independent forward launches are spread over two streams: each launch keeps the device full until its last
 * alignments (a resident grid), so the first blocks of the launch queued on the other stream start in the tail of this one
 * instead of behind it.  begin(): the side stream waits for what `st` holds so far; end(): `st` waits for the side stream. */
bool overlap_launches() {
    static const bool on = [] {
        const char* e = std::getenv("SARLACC_OVERLAP");
        return e ? std::atoi(e) != 0 : true;
    }();
    return on;
}

struct SideStream {
    cudaStream_t aux = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    cudaStream_t begin(cudaStream_t st) {
        if (!overlap_launches()) return st;
        if (!aux) {
            CUDA_CHECK(cudaStreamCreateWithFlags(&aux, cudaStreamNonBlocking));
            CUDA_CHECK(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
            CUDA_CHECK(cudaEventCreateWithFlags(&join, cudaEventDisableTiming));
        }
        CUDA_CHECK(cudaEventRecord(fork, st));
        CUDA_CHECK(cudaStreamWaitEvent(aux, fork, 0));
        return aux;
    }
    void end(cudaStream_t st) {
        if (!aux || !overlap_launches()) return;
        CUDA_CHECK(cudaEventRecord(join, aux));
        CUDA_CHECK(cudaStreamWaitEvent(st, join, 0));
    }
    void release() {
        if (aux) cudaStreamDestroy(aux);
        if (fork) cudaEventDestroy(fork);
        if (join) cudaEventDestroy(join);
        aux = nullptr;
        fork = join = nullptr;
    }
};

struct PairScratch {
    Scratch s[4];
    DevBuf map[2];
    SideStream side;
    DevBuf spec;            /* speculative record writing: strand lists, positions, predictions, ranges, counters */
    DevBuf seeds;           /* 8-mer seed bitmaps of the two adaptors */
    std::string seeds_key;
    bool seeds_usable = false;
    void release() {
        for (auto& x : s) x.release();
        map[0].release();
        map[1].release();
        spec.release();
        seeds.release();
        seeds_key.clear();
        side.release();
    }
};

bool speculate_records() {
    static const bool on = [] {
        const char* e = std::getenv("SARLACC_SPECULATE");
        return e ? std::atoi(e) != 0 : true;
    }();
    return on;
}

/* 8-mer seed table (4096 words, two bits per 8-mer) of the two adaptors' A/C/G/T stretches, for the strand predictor (kernels.h:
 * StrandLists).  Usable only if both adaptors have enough seeds to out-vote chance hits. */
bool prepare_seeds(PairScratch& S, const Plan& p1, const Plan& p2, cudaStream_t st) {
    std::string key;
    for (const Plan* P : {&p1, &p2}) {
        key.append(reinterpret_cast<const char*>(P->refmask.data()), P->refmask.size());
        key += '|';
        key.append(reinterpret_cast<const char*>(P->refkind.data()), P->refkind.size());
        key += '/';
    }
    if (key == S.seeds_key) return S.seeds_usable;
    std::vector<uint32_t> bits(4096, 0);
    int count[2] = {0, 0};
    int which = 0;
    for (const Plan* P : {&p1, &p2}) {
        unsigned code = 0;
        int valid = 0;
        for (int c = 0; c < P->L; ++c) {
            const unsigned m = P->refmask[(size_t)c];
            const bool acgt = P->refkind[(size_t)c] == COL_ACGT && (m == 1 || m == 2 || m == 4 || m == 8);
            const unsigned b = m == 1 ? 0u : (m == 2 ? 1u : (m == 4 ? 2u : 3u));
            code = ((code << 2) | (acgt ? b : 0u)) & 0xFFFFu;
            valid = acgt ? valid + 1 : 0;
            if (valid >= 8) {
                uint32_t& w = bits[code >> 4];
                const unsigned at = (code & 15u) * 2u + (unsigned)which;
                if (!((w >> at) & 1u)) ++count[which];
                w |= 1u << at;
            }
        }
        ++which;
    }
    CUDA_CHECK(cudaStreamSynchronize(st));      /* an earlier launch may still read the previous bitmaps */
    S.seeds.reserve(sizeof(uint32_t) * bits.size());
    CUDA_CHECK(cudaMemcpy(S.seeds.p, bits.data(), sizeof(uint32_t) * bits.size(), cudaMemcpyHostToDevice));
    S.seeds_key = key;
    S.seeds_usable = count[0] >= 6 && count[1] >= 6;
    return S.seeds_usable;
}

struct PairDeviceOut {      /* device pointers: [m] vectors, [nsec][pitch] section matrices */
    uint8_t* reversed = nullptr;
    double* score[2] = {nullptr, nullptr};
    int32_t* start[2] = {nullptr, nullptr};
    int32_t* end[2] = {nullptr, nullptr};
    int32_t* sec_start[2] = {nullptr, nullptr};
    int32_t* sec_width[2] = {nullptr, nullptr};
    long long pitch = 0;
};

/* Scratch bytes per read of one run_pair_device range (four record sets + two column maps). */
size_t pair_scratch_per_read(const Plan& P1, const Plan& P2, int maxlen) {
    auto rec = [&](const Plan& P) -> size_t {
        if (!P.fast) return (size_t)std::max(1, maxlen) * P.L + 4;
        const Geometry g = geometry_for(P, maxlen);
        const size_t wb = (size_t)trace_word_bytes(g.C) + (size_t)trace_hi_bytes(g.C, g.solo);
        return (g.solo ? (size_t)(maxlen + 1) * wb : (size_t)(maxlen + kSkew * g.G) * g.G * wb) + 4;
    };
    return 2 * rec(P1) + 2 * rec(P2) + sizeof(int32_t) * ((size_t)P1.L + P2.L + 2);
}

/* Enqueues on `st` the four forward passes + strand resolution and on `tb` (behind fwd_done) the two tracebacks;
 * tb_done is recorded behind those.  The caller makes `st` wait for the previous user of S before calling. */
const char* run_pair_device(const Plan* const plan[2], const DevPlan* const D[2], PairScratch& S, cudaStream_t st, cudaStream_t tb,
        cudaEvent_t fwd_done, cudaEvent_t tb_done,
        const uint16_t* rows_f, const int32_t* lens_f, int stride_f, const uint16_t* rows_b, const int32_t* lens_b, int stride_b,
        long long m, int maxlen, const int32_t* width, double* tmp_scores, const PairDeviceOut& out, int sms, FwdTimer* timer = nullptr)
{
    if (m <= 0) return "";
    Range nvtx("sarlacc: both adaptors x both windows");
    /* Speculative record writing (kernels.h: StrandLists): only for the wavefront kernels' row-pair form, whose launches
     * can take a list of reads with device-side bounds. */
    /* below ~40 000 reads the extra launches cost more than the records save (28 416 reads: 4.88 vs 4.63 ms per call,
     * 56 832: 7.87 vs 8.01); SARLACC_SPEC_MIN moves the bound (tests: 512) */
    const char* spec_min_env = std::getenv("SARLACC_SPEC_MIN");
    const long long spec_min = spec_min_env && std::atoll(spec_min_env) > 0 ? std::atoll(spec_min_env) : 40000;
    bool spec = speculate_records() && dynamic_distribution() && m >= spec_min && m < (1LL << 31);
    for (int a = 0; a < 2 && spec; ++a) spec = plan[a]->fast && geometry_for(*plan[a], maxlen).pair != 0;
    if (spec) spec = prepare_seeds(S, *plan[0], *plan[1], st);
    StrandLists L;
    std::memset(&L, 0, sizeof(L));
    unsigned long long* next = nullptr;
    if (spec) {
        const size_t mm = (size_t)m;
        auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
        const size_t o_lists = 0, o_pos = o_lists + al(sizeof(int32_t) * 4 * mm), o_pred = o_pos + al(sizeof(int32_t) * 2 * mm),
                     o_ranges = o_pred + al(mm), o_next = o_ranges + 256, total = o_next + 256;
        S.spec.reserve(total);
        uint8_t* base = S.spec.as<uint8_t>();
        L.list_fwd = reinterpret_cast<int32_t*>(base + o_lists);
        L.list_rev = L.list_fwd + mm;
        L.list_sfwd = L.list_rev + mm;
        L.list_srev = L.list_sfwd + mm;
        L.pos_fwd = reinterpret_cast<int32_t*>(base + o_pos);
        L.pos_rev = L.pos_fwd + mm;
        L.predicted = base + o_pred;
        L.ranges = reinterpret_cast<int32_t*>(base + o_ranges);
        next = reinterpret_cast<unsigned long long*>(base + o_next);
        CUDA_CHECK(cudaMemsetAsync(base + o_ranges, 0, 512, st));      /* ranges and the twelve launch counters */
        ClassifyArgs CA;
        std::memset(&CA, 0, sizeof(CA));
        CA.rows_front = rows_f;
        CA.rows_back = rows_b;
        CA.lens_front = lens_f;
        CA.lens_back = lens_b;
        CA.n = m;
        CA.stride = stride_f;
        CA.seeds = S.seeds.as<uint32_t>();
        CA.margin = 3;
        {
            /* the predictor reads the first bases of each window only: an adaptor further in than that leaves the read
             * "unsure" (records on both strands, as without speculation) */
            static const int scan = [] {
                const char* e = std::getenv("SARLACC_SPEC_SCAN");
                return e && std::atoi(e) > 0 ? std::atoi(e) : 128;
            }();
            CA.scan = scan;
        }
        /* test switch (read per call): 1 turns every prediction round (all reads take the re-run path unless unsure),
         * 2 makes every read unsure (both strands with records) */
        const char* tmode = std::getenv("SARLACC_SPEC_TEST");
        CA.test_mode = tmode ? std::atoi(tmode) : 0;
        CA.L = L;
        if (stride_f != stride_b) spec = false;
        else {
            launch_classify_strands(CA, st);
            g_launches += 2;
        }
    }
    FwdRec rec[4];
    cudaStream_t side = S.side.begin(st);
    for (int r = 0; r < 4; ++r) {
        const int a = (r == 0 || r == 2) ? 0 : 1;          /* adaptor */
        const bool on_front = (r == 0 || r == 3);           /* window set */
        const bool rev_strand = r >= 2;                     /* (adaptor1, back) and (adaptor2, front) are the reverse strand's passes */
        Outputs dev;
        dev.score = tmp_scores + (size_t)r * m;
        if (!spec) {
            rec[r] = forward_once(*plan[a], *D[a], S.s[r], rev_strand ? side : st, on_front ? rows_f : rows_b, on_front ? lens_f : lens_b, m,
                                  on_front ? stride_f : stride_b, maxlen, true, dev, sms, r == 0 ? timer : nullptr);
            continue;
        }
        FwdOrder with_records, score_only;
        with_records.index = rev_strand ? L.list_rev : L.list_fwd;
        with_records.range = L.ranges + (rev_strand ? 2 : 0);
        with_records.next = next + 2 * r;
        score_only.index = rev_strand ? L.list_srev : L.list_sfwd;
        score_only.range = L.ranges + (rev_strand ? 6 : 4);
        score_only.next = next + 2 * r + 1;
        score_only.fill_empty = false;
        rec[r] = forward_once(*plan[a], *D[a], S.s[r], st, on_front ? rows_f : rows_b, on_front ? lens_f : lens_b, m,
                              on_front ? stride_f : stride_b, maxlen, true, dev, sms, r == 0 ? timer : nullptr, nullptr, &with_records);
        forward_once(*plan[a], *D[a], S.s[r], side, on_front ? rows_f : rows_b, on_front ? lens_f : lens_b, m,
                     on_front ? stride_f : stride_b, maxlen, false, dev, sms, nullptr, nullptr, &score_only);
    }
    S.side.end(st);
    StrandArgs SA;
    std::memset(&SA, 0, sizeof(SA));
    SA.n = m;
    SA.a1_front = tmp_scores;
    SA.a2_back = tmp_scores + (size_t)m;
    SA.a1_back = tmp_scores + (size_t)2 * m;
    SA.a2_front = tmp_scores + (size_t)3 * m;
    SA.reversed = out.reversed;
    SA.score1 = out.score[0];
    SA.score2 = out.score[1];
    SA.strand_score = nullptr;
    if (spec) {
        SA.predicted = L.predicted;
        SA.list_fwd = L.list_fwd;
        SA.list_rev = L.list_rev;
        SA.pos_fwd = L.pos_fwd;
        SA.pos_rev = L.pos_rev;
        SA.ranges = L.ranges;
    }
    launch_resolve_strand(SA, st);
    g_launches += 1;
    if (spec) {
        /* the reads whose kept strand was scored without records: a second pass with records, appended to the same record
         * sets (positions behind the predicted ones); normally a handful, so a small grid */
        for (int r = 0; r < 4; ++r) {
            const int a = (r == 0 || r == 2) ? 0 : 1;
            const bool on_front = (r == 0 || r == 3);
            const bool rev_strand = r >= 2;
            Outputs dev;
            dev.score = tmp_scores + (size_t)r * m;
            FwdOrder redo;
            redo.index = rev_strand ? L.list_rev : L.list_fwd;
            redo.range = L.ranges + (rev_strand ? 10 : 8);
            redo.next = next + 8 + r;
            redo.fill_empty = false;
            forward_once(*plan[a], *D[a], S.s[r], st, on_front ? rows_f : rows_b, on_front ? lens_f : lens_b, m,
                         on_front ? stride_f : stride_b, maxlen, true, dev, sms, nullptr, nullptr, &redo);
        }
    }
    if (spec && std::getenv("SARLACC_DEBUG_SPEC")) {
        int32_t h[12];
        CUDA_CHECK(cudaStreamSynchronize(st));
        CUDA_CHECK(cudaMemcpy(h, L.ranges, sizeof(h), cudaMemcpyDeviceToHost));
        std::fprintf(stderr, "[sarlacc] speculation: %lld reads, records fwd %d rev %d, score-only fwd %d rev %d, re-run fwd %d rev %d\n",
                     m, h[1], h[3], h[5], h[7], h[9] - h[8], h[11] - h[10]);
    }
    CUDA_CHECK(cudaEventRecord(fwd_done, st));
    CUDA_CHECK(cudaStreamWaitEvent(tb, fwd_done, 0));
    for (int a = 0; a < 2; ++a) {
        const FwdRec& fwd = rec[a == 0 ? 0 : 1];     /* forward strand: (adaptor1, front) / (adaptor2, back) */
        const FwdRec& rev = rec[a == 0 ? 2 : 3];     /* reverse strand: (adaptor1, back) / (adaptor2, front) */
        TraceArgs T = trace_args_for(*plan[a], *D[a], fwd);
        T.sel = out.reversed;
        T.lens2 = rev.A.lens;
        T.flags2 = rev.A.flags;
        T.flags_hi2 = rev.A.flags_hi;
        T.endrow2 = rev.A.endrow;
        if (spec) {
            T.pos = L.pos_fwd;
            T.pos2 = L.pos_rev;
        }
        S.map[a].reserve(sizeof(int32_t) * ((size_t)plan[a]->L + 1) * (size_t)m);
        T.map = S.map[a].as<int32_t>();
        T.start = out.start[a];
        T.end = out.end[a];
        T.sec_start = out.sec_start[a];
        T.sec_width = out.sec_width[a];
        T.out_pitch = out.pitch;
        T.width = a == 1 ? width : nullptr;
        launch_traceback(T, tb);
        g_launches += 1;
    }
    CUDA_CHECK(cudaEventRecord(tb_done, tb));
    CUDA_CHECK(cudaGetLastError());
    g_last_kernel = rec[0].name;
    return rec[0].name;
}

}  // namespace

/* The [nref][n] and [nsec][n] output matrices need the TOTAL n as their row pitch when a run is split
 * into sub-chunks.  Rather than thread a pitch through every kernel, runs that need sub-chunking use
 * per-sub-chunk staging: see Job below, which always hands run_device a range it owns entirely. */

namespace {

int device_sm_count(int dev) {
    int sms = 0;
    CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    return sms;
}

enum Mode { MODE_SCORE_LOCAL = 0, MODE_TRACE_LOCAL = 1, MODE_SCORE_GLOBAL = 2, MODE_OPS_GLOBAL = 3, MODE_MULTI_GLOBAL = 4 };

struct HostOutputs {
    double* score = nullptr;       /* [nref][n] (nref == 1 unless MODE_MULTI with all_scores) */
    int32_t* start = nullptr;
    int32_t* end = nullptr;
    int32_t* sec_start = nullptr;
    int32_t* sec_width = nullptr;
    int32_t* best_id = nullptr;
    double* best = nullptr;
    double* next_best = nullptr;
    /* MODE_OPS_GLOBAL: general_align post-processing on the host needs ops per read */
    std::vector<std::vector<uint8_t> >* ops = nullptr;
};

/* Raw staging of one window set of one chunk: the windows' sequence and quality bytes go to the device as they are and
 * are packed there (kernels.cu: pack_rows_kernel), so the host only moves bytes.  On a box with few cores per GPU (the
 * 8-GPU node has 4) the host packer, not the device, set the end-to-end rate.
 *   - CSR input, whole entries (no window cutting), pools in pinned host memory: no staging at all, the chunk's byte
 *     range is DMA'd straight from the caller's pool;
 *   - otherwise the windows are gathered into pinned staging (one memcpy per window, or one per thread range when the
 *     chunk is a contiguous run of a CSR pool). */
struct RawStage {
    PinBuf h_seq, h_qual, h_soff, h_qoff, h_bad;
    DevBuf d_seq, d_qual, d_soff, d_qoff, d_bad;
    bool bad_pending = false;
    void release() {
        h_seq.release(); h_qual.release(); h_soff.release(); h_qoff.release(); h_bad.release();
        d_seq.release(); d_qual.release(); d_soff.release(); d_qoff.release(); d_bad.release();
    }
    /* after the slot's `done` event: index (within the chunk) of the first window with a quality below the offset, or -1 */
    long long first_bad() {
        if (!bad_pending) return -1;
        bad_pending = false;
        const long long v = *h_bad.as<long long>();
        return v == std::numeric_limits<long long>::max() ? -1 : v;
    }
};

/* Copy into pinned staging with non-temporal stores: the destination is read next by the DMA engine, not by a core, so
 * there is no point in pulling its lines into the cache first (a thread's piece, ~2 MB, is below the size at which
 * memcpy switches to such stores by itself). */
#ifdef SARLACC_HAVE_AVX2_PACK
__attribute__((target("avx2"))) static void copy_nt_avx2(uint8_t* dst, const uint8_t* src, size_t n) {
    size_t head = (32 - (reinterpret_cast<uintptr_t>(dst) & 31)) & 31;
    if (head > n) head = n;
    std::memcpy(dst, src, head);
    size_t i = head;
    for (; i + 128 <= n; i += 128) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 64));
        const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), a);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 96), d);
    }
    _mm_sfence();
    std::memcpy(dst + i, src + i, n - i);
}
#endif

void copy_to_staging(uint8_t* dst, const uint8_t* src, size_t n) {
#ifdef SARLACC_HAVE_AVX2_PACK
    static const bool nt = have_avx2() && std::getenv("SARLACC_NO_NT_COPY") == nullptr;
    if (nt && n >= 4096) {
        copy_nt_avx2(dst, src, n);
        return;
    }
#endif
    std::memcpy(dst, src, n);
}

/* Sequence bytes as 4-bit one-hot base codes, two per byte (base k of the range in nibble k & 1 of byte k >> 1): what a
 * job that waits for its uploads sends instead of the bytes themselves (see stage_and_pack).  The codes are the
 * forward packer table's (PackTables::base): four byte values map to 1, 2, 4, 8, everything else to 0, so the
 * device packer gets exactly what it would have looked up itself. */
#ifdef SARLACC_HAVE_AVX2_PACK
struct NibbleConsts { __m256i c0, c1, c2, c3, b1, b2, b4, b8; };
__attribute__((target("avx2"))) static inline __m256i nibble_codes_avx2(__m256i x, const NibbleConsts& K) {
    __m256i k = _mm256_and_si256(_mm256_cmpeq_epi8(x, K.c0), K.b1);
    k = _mm256_or_si256(k, _mm256_and_si256(_mm256_cmpeq_epi8(x, K.c1), K.b2));
    k = _mm256_or_si256(k, _mm256_and_si256(_mm256_cmpeq_epi8(x, K.c2), K.b4));
    return _mm256_or_si256(k, _mm256_and_si256(_mm256_cmpeq_epi8(x, K.c3), K.b8));
}
__attribute__((target("avx2"))) static size_t nibble_pack_avx2(const uint8_t* src, size_t n, uint8_t* dst, const uint8_t code[4]) {
    NibbleConsts K;
    K.c0 = _mm256_set1_epi8((char)code[0]); K.c1 = _mm256_set1_epi8((char)code[1]);
    K.c2 = _mm256_set1_epi8((char)code[2]); K.c3 = _mm256_set1_epi8((char)code[3]);
    K.b1 = _mm256_set1_epi8(1); K.b2 = _mm256_set1_epi8(2); K.b4 = _mm256_set1_epi8(4); K.b8 = _mm256_set1_epi8(8);
    const __m256i mul = _mm256_set1_epi16(0x1001);      /* even byte x 1 + odd byte x 16 */
    size_t i = 0;
    for (; i + 64 <= n; i += 64) {
        const __m256i w0 = _mm256_maddubs_epi16(nibble_codes_avx2(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i)), K), mul);
        const __m256i w1 = _mm256_maddubs_epi16(nibble_codes_avx2(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32)), K), mul);
        const __m256i p = _mm256_permute4x64_epi64(_mm256_packus_epi16(w0, w1), 0xD8);
        _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + i / 2), p);
    }
    return i;
}
#endif

/* dst[(n + 1) / 2] <- nibbles of src[0, n); threads split the range at multiples of 64 bases */
void nibble_pack(const uint8_t* src, size_t n, uint8_t* dst, const PackTables& T, int nthreads) {
    const int64_t blocks = (int64_t)((n + 63) / 64);
    parallel_for(0, blocks, nthreads, [&](int64_t a, int64_t b, int) {
        size_t lo = (size_t)a * 64, hi = std::min(n, (size_t)b * 64);
        size_t i = lo;
#ifdef SARLACC_HAVE_AVX2_PACK
        if (have_avx2()) i += nibble_pack_avx2(src + lo, hi - lo, dst + lo / 2, T.code);
#endif
        for (; i + 1 < hi; i += 2) dst[i / 2] = (uint8_t)(T.base[src[i]] | (T.base[src[i + 1]] << 4));
        if (i < hi) dst[i / 2] = T.base[src[i]];
    });
}

bool pointer_is_pinned(const void* p) {
    if (!p) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

/* Enqueues on `st`: upload of the raw bytes of windows [c0, c1) of V + the device packer writing d_rows. */
struct UploadBytes { size_t sent = 0, raw = 0; };      /* bytes a chunk's window sets put on the link / would have put as plain bytes */

/* The device packer of a staged window set and the read-back of its error flag, to be enqueued behind the uploads. */
struct PendingPack {
    PackArgs args;
    RawStage* stage = nullptr;
    bool check_qual = false;
    bool valid = false;
};

void launch_pending_pack(PendingPack& P, cudaStream_t st) {
    if (!P.valid) return;
    launch_pack_rows(P.args, st);
    g_launches += 1;
    CUDA_CHECK(cudaMemcpyAsync(P.stage->h_bad.p, P.stage->d_bad.p, sizeof(long long), cudaMemcpyDeviceToHost, st));
    P.stage->bad_pending = P.check_qual;
    P.valid = false;
}

/* With `defer` the packer is not launched here but handed back: a caller with two window sets enqueues BOTH uploads
 * first and the two packers after them.  The packer has to wait for free SMs while a forward kernel of the previous
 * chunk runs, and everything behind it on the stream waits with it -- the second window set's upload included, which
 * left the copy engine idle for a kernel's length per chunk (N = 8: 11.6 GB/s per rank against 23 GB/s available,
 * profiles/r02_history.md). */
/* The host half of stage_and_pack: offsets and, where needed, the staged (or 4-bit coded) bytes of windows [c0, c1). */
struct Staged {
    const uint8_t* src_seq = nullptr;
    const uint8_t* src_qual = nullptr;
    size_t nseq = 0, nqual = 0;
    bool seq4 = false;
    long long m = 0;
};

Staged stage_host(const ReadView& V, int64_t c0, int64_t c1, const PackTables& T, const int32_t* h_lens, bool pools_pinned,
        RawStage& R, int nthreads, bool nibbles)
{
    Staged G;
    const long long m = c1 - c0;
    G.m = m;
    if (m <= 0) return G;
    Range nvtx("sarlacc: stage windows");
    R.h_soff.reserve(sizeof(long long) * (size_t)m);
    R.h_qoff.reserve(sizeof(long long) * (size_t)m);
    long long* soff = R.h_soff.as<long long>();
    long long* qoff = R.h_qoff.as<long long>();
    const sarlacc_reads* S = V.R;
    const bool csr_whole = !S->seq && V.tol == 0;
    const uint8_t* src_seq = nullptr;
    const uint8_t* src_qual = nullptr;
    size_t nseq = 0, nqual = 0;
    bool seq4 = false;
    if (csr_whole) {
        /* the chunk is one contiguous byte range of each pool; a window's offsets are its entry's */
        const int64_t s0 = S->seq_off[c0], q0 = S->qual_off[c0];
        for (long long i = 0; i < m; ++i) {
            soff[i] = S->seq_off[c0 + i] - s0;
            qoff[i] = S->qual_off[c0 + i] - q0;
        }
        nseq = (size_t)(S->seq_off[c1] - s0);
        nqual = (size_t)(S->qual_off[c1] - q0);
        if (nibbles) {
            /* SARLACC_PACK_SEQ=1 (see PairJob): the bases go up as 4-bit codes -- half their bytes for one pass of the
             * host over them; qualities as they are */
            seq4 = true;
            R.h_seq.reserve((nseq + 1) / 2 + 64);
            nibble_pack(S->seq_pool + s0, nseq, R.h_seq.as<uint8_t>(), T, nthreads);
            src_seq = R.h_seq.as<uint8_t>();
            if (pools_pinned) {
                src_qual = S->qual_pool + q0;
            } else {
                R.h_qual.reserve(nqual);
                uint8_t* hq = R.h_qual.as<uint8_t>();
                parallel_for(0, m, nthreads, [&](int64_t a, int64_t b, int) {
                    copy_to_staging(hq + qoff[a], S->qual_pool + q0 + qoff[a], (size_t)(S->qual_off[c0 + b] - S->qual_off[c0 + a]));
                });
                src_qual = hq;
            }
        } else if (pools_pinned) {
            src_seq = S->seq_pool + s0;
            src_qual = S->qual_pool + q0;
        } else {
            R.h_seq.reserve(nseq);
            R.h_qual.reserve(nqual);
            uint8_t* hs = R.h_seq.as<uint8_t>();
            uint8_t* hq = R.h_qual.as<uint8_t>();
            parallel_for(0, m, nthreads, [&](int64_t a, int64_t b, int) {
                copy_to_staging(hs + soff[a], S->seq_pool + s0 + soff[a], (size_t)(S->seq_off[c0 + b] - S->seq_off[c0 + a]));
                copy_to_staging(hq + qoff[a], S->qual_pool + q0 + qoff[a], (size_t)(S->qual_off[c0 + b] - S->qual_off[c0 + a]));
            });
            src_seq = hs;
            src_qual = hq;
        }
    } else {
        /* views and cut windows: gather the windows back to back (same offsets for bases and qualities) */
        long long at = 0;
        for (long long i = 0; i < m; ++i) {
            soff[i] = qoff[i] = at;
            at += h_lens[i];
        }
        nseq = nqual = (size_t)at;
        R.h_seq.reserve(nseq);
        R.h_qual.reserve(nqual);
        uint8_t* hs = R.h_seq.as<uint8_t>();
        uint8_t* hq = R.h_qual.as<uint8_t>();
        parallel_for(0, m, nthreads, [&](int64_t a, int64_t b, int) {
            for (int64_t i = a; i < b; ++i) {
                const int len = h_lens[i];
                if (len <= 0) continue;
                std::memcpy(hs + soff[i], V.seq(c0 + i), (size_t)len);
                std::memcpy(hq + qoff[i], V.qual(c0 + i), (size_t)len);
            }
        });
        src_seq = hs;
        src_qual = hq;
    }
    G.src_seq = src_seq;
    G.src_qual = src_qual;
    G.nseq = nseq;
    G.nqual = nqual;
    G.seq4 = seq4;
    return G;
}

/* The device half: uploads of what stage_host prepared + the device packer (handed back with `defer`). */
void stage_enqueue(const Staged& G, const ReadView& V, const PackTables& T, const int32_t* d_lens, int stride, uint16_t* d_rows,
        bool check_qual, RawStage& R, cudaStream_t st, PendingPack* defer = nullptr, UploadBytes* bytes = nullptr)
{
    const long long m = G.m;
    if (m <= 0) return;
    Range nvtx("sarlacc: H2D + device pack");
    const uint8_t* src_seq = G.src_seq;
    const uint8_t* src_qual = G.src_qual;
    const size_t nseq = G.nseq, nqual = G.nqual;
    const bool seq4 = G.seq4;
    const long long* soff = R.h_soff.as<long long>();
    const long long* qoff = R.h_qoff.as<long long>();
    R.d_seq.reserve(nseq);
    R.d_qual.reserve(nqual);
    R.d_soff.reserve(sizeof(long long) * (size_t)m);
    R.d_qoff.reserve(sizeof(long long) * (size_t)m);
    R.d_bad.reserve(sizeof(long long));
    R.h_bad.reserve(sizeof(long long));
    const size_t seq_bytes = seq4 ? (nseq + 1) / 2 : nseq;
    if (bytes) {
        bytes->sent += seq_bytes + nqual + 2 * sizeof(long long) * (size_t)m;
        bytes->raw += nseq + nqual + 2 * sizeof(long long) * (size_t)m;
    }
    if (nseq) CUDA_CHECK(cudaMemcpyAsync(R.d_seq.p, src_seq, seq_bytes, cudaMemcpyHostToDevice, st));
    if (nqual) CUDA_CHECK(cudaMemcpyAsync(R.d_qual.p, src_qual, nqual, cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaMemcpyAsync(R.d_soff.p, soff, sizeof(long long) * (size_t)m, cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaMemcpyAsync(R.d_qoff.p, qoff, sizeof(long long) * (size_t)m, cudaMemcpyHostToDevice, st));
    *R.h_bad.as<long long>() = std::numeric_limits<long long>::max();
    CUDA_CHECK(cudaMemcpyAsync(R.d_bad.p, R.h_bad.p, sizeof(long long), cudaMemcpyHostToDevice, st));
    PackArgs A;
    A.seq = R.d_seq.as<uint8_t>();
    A.qual = R.d_qual.as<uint8_t>();
    A.soff = R.d_soff.as<long long>();
    A.qoff = R.d_qoff.as<long long>();
    A.lens = d_lens;
    A.n = m;
    A.stride = stride;
    A.back = V.back ? 1 : 0;
    A.seq4 = seq4 ? 1 : 0;
    A.rows = d_rows;
    A.first_bad = check_qual ? R.d_bad.as<long long>() : nullptr;
    std::memcpy(A.base, T.base, 256);
    std::memcpy(A.base_rc, T.base_rc, 256);
    std::memcpy(A.qidx, T.qidx, 512);
    PendingPack own;
    PendingPack& P = defer ? *defer : own;
    P.args = A;
    P.stage = &R;
    P.check_qual = check_qual;
    P.valid = true;
    if (!defer) launch_pending_pack(P, st);
}

void stage_and_pack(const ReadView& V, int64_t c0, int64_t c1, const PackTables& T, const int32_t* h_lens, const int32_t* d_lens,
        int stride, uint16_t* d_rows, bool check_qual, bool pools_pinned, RawStage& R, cudaStream_t st, int nthreads,
        PendingPack* defer = nullptr, bool nibbles = false, UploadBytes* bytes = nullptr)
{
    const Staged G = stage_host(V, c0, c1, T, h_lens, pools_pinned, R, nthreads, nibbles);
    stage_enqueue(G, V, T, d_lens, stride, d_rows, check_qual, R, st, defer, bytes);
}

/* One pipeline slot: pinned staging + device buffers for a chunk of reads. */
struct Slot {
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr;
    cudaEvent_t t_begin = nullptr, t_h2d = nullptr, t_end = nullptr;   /* SARLACC_DEBUG_TIMING: per-chunk device phases */
    /* recorded behind the chunk's alignment kernels.  Chunks whose kernels must not start before the previous chunk's
     * have finished wait for it: those with tracebacks in the single-adaptor jobs (round 1 measured two static-grid
     * forward kernels sharing the device at three chunks in the time of 3.4).  Score-only chunks and the both-ends job
     * do not wait (SARLACC_GATE_CHUNKS=1 makes them): their launches are resident grids fed from a device counter, and
     * the next chunk's blocks simply start as this chunk's blocks retire. */
    cudaEvent_t gate = nullptr;
    PinBuf h_rows, h_lens, h_out, h_order;
    DevBuf d_rows, d_lens, d_out, d_order;     /* *_order: the chunk's reads by length (barcode-length reads) */
    PinBuf h_rows2, h_lens2, h_width;     /* second window set + read widths of the fused both-ends entry */
    DevBuf d_rows2, d_lens2, d_width, d_tmp;
    Scratch scratch;
    cudaStream_t tb = nullptr;    /* tracebacks of the fused entry */
    cudaEvent_t fwd_ev[4] = {nullptr, nullptr, nullptr, nullptr}, tb_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    RawStage raw, raw2;     /* raw bytes of the (first, second) window set on their way to the device packer */
    long long n = 0;        /* reads in flight */
    int64_t lo = 0;
    bool busy = false;
    UploadBytes up;         /* of the chunk in flight (both-ends job) */
    void init() {
        CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        CUDA_CHECK(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&gate, cudaEventDisableTiming));
        CUDA_CHECK(cudaStreamCreateWithFlags(&tb, cudaStreamNonBlocking));
        for (int k = 0; k < 4; ++k) {
            CUDA_CHECK(cudaEventCreateWithFlags(&fwd_ev[k], cudaEventDisableTiming));
            CUDA_CHECK(cudaEventCreateWithFlags(&tb_ev[k], cudaEventDisableTiming));
        }
        CUDA_CHECK(cudaEventCreate(&t_begin));
        CUDA_CHECK(cudaEventCreate(&t_h2d));
        CUDA_CHECK(cudaEventCreate(&t_end));
    }
    void destroy() {
        h_rows.release(); h_lens.release(); h_out.release(); h_order.release();
        d_rows.release(); d_lens.release(); d_out.release(); d_order.release();
        h_rows2.release(); h_lens2.release(); h_width.release();
        d_rows2.release(); d_lens2.release(); d_width.release(); d_tmp.release();
        scratch.release();
        raw.release();
        raw2.release();
        if (done) cudaEventDestroy(done);
        if (gate) cudaEventDestroy(gate);
        gate = nullptr;
        for (int k = 0; k < 4; ++k) {
            if (fwd_ev[k]) cudaEventDestroy(fwd_ev[k]);
            if (tb_ev[k]) cudaEventDestroy(tb_ev[k]);
            fwd_ev[k] = tb_ev[k] = nullptr;
        }
        if (tb) cudaStreamDestroy(tb);
        tb = nullptr;
        if (t_begin) cudaEventDestroy(t_begin);
        if (t_h2d) cudaEventDestroy(t_h2d);
        if (t_end) cudaEventDestroy(t_end);
        t_begin = t_h2d = t_end = nullptr;
        if (st) cudaStreamDestroy(st);
        done = nullptr;
        st = nullptr;
    }
};

/* Layout of a slot's output block (same on host and device). */
struct OutLayout {
    size_t o_score = 0, o_start = 0, o_end = 0, o_ss = 0, o_sw = 0, o_bid = 0, o_best = 0, o_next = 0, o_nops = 0, o_ops = 0, total = 0;
    long long ops_stride = 0;
};

OutLayout make_layout(Mode mode, long long n, int nref_scores, int nsec, int maxlen, int L) {
    OutLayout o;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t at = 0;
    o.o_score = at; at += al(sizeof(double) * (size_t)n * std::max(1, nref_scores));
    if (mode == MODE_TRACE_LOCAL) {
        o.o_start = at; at += al(sizeof(int32_t) * (size_t)n);
        o.o_end = at; at += al(sizeof(int32_t) * (size_t)n);
        o.o_ss = at; at += al(sizeof(int32_t) * (size_t)n * std::max(1, nsec));
        o.o_sw = at; at += al(sizeof(int32_t) * (size_t)n * std::max(1, nsec));
    }
    if (mode == MODE_MULTI_GLOBAL) {
        o.o_bid = at; at += al(sizeof(int32_t) * (size_t)n);
        o.o_best = at; at += al(sizeof(double) * (size_t)n);
        o.o_next = at; at += al(sizeof(double) * (size_t)n);
    }
    if (mode == MODE_OPS_GLOBAL) {
        o.ops_stride = ((long long)maxlen + L + 16) & ~15LL;
        o.o_nops = at; at += al(sizeof(int32_t) * (size_t)n);
        o.o_ops = at; at += al((size_t)o.ops_stride * (size_t)n);
    }
    o.total = at;
    return o;
}

/* Per-device resources kept between host-buffer calls.  One call at a time per device (the R boundary is
 * single-threaded; concurrent callers serialise on the device's mutex). */
constexpr int kMaxSlots = 8;
/* Pipeline depth of the host-buffer entries: three slots -- while the device runs chunk k and chunk k+1's upload is in
 * flight, the host stages chunk k+2.  Deeper pipelines were tried for the 8-GPU case, where all ranks upload at once and
 * a rank gets 23-35 GB/s instead of 55 (tools/h2d_probe.py): 5 and 7 slots were no faster (75 / 62 ms per 1 M reads on
 * the slow / fast half of the GPUs with 3 slots, 78 / 67 with 5, 72 / 68 with 7; profiles/r02_history.md), so the limit
 * there is the upload rate itself, not pauses of the copy engine.  SARLACC_SLOTS (2..8) overrides. */
int slot_count() {
    static const int n = [] {
        const char* e = std::getenv("SARLACC_SLOTS");
        const int v = e ? std::atoi(e) : 3;
        return v < 2 ? 2 : (v > kMaxSlots ? kMaxSlots : v);
    }();
    return n;
}
#define kSlots (slot_count())

struct DeviceCache {
    std::mutex busy;
    bool ready = false;
    /* three slots: while the device runs chunk k and chunk k+1's upload is in flight, the host packs chunk k+2
     * (with two, the upload of k+1 could only start after k-1 had finished and k+1 been packed: the device idled
     * ~2 ms of every 7.5 ms chunk period) */
    Slot slots[kMaxSlots];
    DevPlan plan;
    /* the fused both-ends entry: four record sets per chunk, two groups used alternately by consecutive chunks (the
     * tracebacks of chunk k run beside the forward passes of chunk k+1) -- shared by all slots, a slot's own memory is
     * its raw bytes, rows and result columns */
    PairScratch pair[2];
    static DeviceCache& acquire(int device) {
        static std::mutex table_mutex;
        static std::vector<std::unique_ptr<DeviceCache> > table;
        DeviceCache* c = nullptr;
        {
            std::lock_guard<std::mutex> lock(table_mutex);
            if ((int)table.size() <= device) table.resize(device + 1);
            if (!table[device]) table[device].reset(new DeviceCache());
            c = table[device].get();
        }
        c->busy.lock();
        if (!c->ready) {
            try {
                for (int k = 0; k < kSlots; ++k) c->slots[k].init();
            } catch (...) {
                c->busy.unlock();
                throw;
            }
            c->ready = true;
        }
        for (int k = 0; k < kSlots; ++k) c->slots[k].busy = false;
        return *c;
    }
    void release() {
        /* make sure nothing of this call is still in flight before another call reuses the buffers */
        if (ready) {
            for (int k = 0; k < kSlots; ++k) cudaStreamSynchronize(slots[k].st);
        }
        busy.unlock();
    }
};

/* Runs [lo,hi) of the reads on one device with a two-slot pipeline: pack chunk k+1 on the host while the
 * device works on chunk k. */
struct DeviceJob {
    int device = 0;
    int64_t lo = 0, hi = 0;
    const sarlacc_reads* reads = nullptr;
    const Plan* plan = nullptr;
    Mode mode = MODE_SCORE_LOCAL;
    int64_t n_total = 0;
    HostOutputs out;
    bool want_all_scores = false;
    int nthreads = 1;
    FirstError err;
    std::string cuda_error;

    void run() {
        try {
            run_inner();
        } catch (CudaError& e) {
            cuda_error = e.msg;
        } catch (std::exception& e) {
            cuda_error = std::string("internal error: ") + e.what();
        }
    }

    void run_inner() {
        const Plan& P = *plan;
        CUDA_CHECK(cudaSetDevice(device));
        const int sms = device_sm_count(device);
        ReadView V{reads};
        PackTables PT;
        build_pack_tables(PT, reads->seq_encoding, *P.enc);
        const bool trace = (mode == MODE_TRACE_LOCAL || mode == MODE_OPS_GLOBAL);
        const int nsec = (int)P.sec_starts.size();
        const int nref_scores = (mode == MODE_MULTI_GLOBAL) ? (want_all_scores ? P.nref : 0) : 1;

        /* Streams, pinned staging, device buffers and scratch are cached per device across calls
         * (cudaMallocHost / cudaMalloc / cudaFree of hundreds of MB per call would dominate short calls). */
        DeviceCache& cache = DeviceCache::acquire(device);
        struct Release {
            DeviceCache& c;
            ~Release() { c.release(); }
        } release{cache};
        Slot* slots = cache.slots;
        DevPlan& D = cache.plan;
        D.upload(P, slots[0].st);

        /* chunk size: bounded so that staging stays modest and the trace scratch fits its budget; a whole number of
         * grid-fulls of alignments (no tail round), and a smaller first chunk so that the device starts early */
        const long long groups = plan_groups(P, trace);
        long long chunk = chunk_for(1 << 17, groups, 0);
        long long first_chunk = (hi - lo > chunk && groups > 0) ? std::max<long long>(1, whole_rounds(chunk / 4, groups)) : chunk;
        const char* ce = std::getenv("SARLACC_CHUNK");
        if (ce && std::atoll(ce) > 0) first_chunk = chunk = std::atoll(ce);

        OutLayout lay[kMaxSlots];
        int which = 0;
        auto drain = [&](Slot& s, const OutLayout& o) {
            if (!s.busy) return;
            Range nvtx("sarlacc: wait for chunk + copy out");
            CUDA_CHECK(cudaEventSynchronize(s.done));
            {
                const long long fb = s.raw.first_bad();      /* quality below the offset, found by the device packer */
                if (fb >= 0) err.offer(s.lo + fb, ERR_QUAL);
            }
            const uint8_t* h = s.h_out.as<uint8_t>();
            const long long m = s.n;
            const int64_t g0 = s.lo;   /* global index of the slot's first read */
            if (mode == MODE_MULTI_GLOBAL) {
                std::memcpy(out.best_id + g0, h + o.o_bid, sizeof(int32_t) * m);
                std::memcpy(out.best + g0, h + o.o_best, sizeof(double) * m);
                std::memcpy(out.next_best + g0, h + o.o_next, sizeof(double) * m);
                if (want_all_scores) {
                    for (int b = 0; b < P.nref; ++b) {
                        std::memcpy(out.score + (size_t)b * n_total + g0, h + o.o_score + sizeof(double) * (size_t)b * m, sizeof(double) * m);
                    }
                }
            } else {
                std::memcpy(out.score + g0, h + o.o_score, sizeof(double) * m);
            }
            if (mode == MODE_TRACE_LOCAL) {
                std::memcpy(out.start + g0, h + o.o_start, sizeof(int32_t) * m);
                std::memcpy(out.end + g0, h + o.o_end, sizeof(int32_t) * m);
                for (int sct = 0; sct < nsec; ++sct) {
                    std::memcpy(out.sec_start + (size_t)sct * n_total + g0, h + o.o_ss + sizeof(int32_t) * (size_t)sct * m, sizeof(int32_t) * m);
                    std::memcpy(out.sec_width + (size_t)sct * n_total + g0, h + o.o_sw + sizeof(int32_t) * (size_t)sct * m, sizeof(int32_t) * m);
                }
            }
            if (mode == MODE_OPS_GLOBAL) {
                const int32_t* nops = reinterpret_cast<const int32_t*>(h + o.o_nops);
                for (long long i = 0; i < m; ++i) {
                    const uint8_t* p = h + o.o_ops + (size_t)i * o.ops_stride;
                    (*out.ops)[g0 + i].assign(p, p + nops[i]);
                }
            }
            s.busy = false;
        };

        cudaEvent_t prev_upload = nullptr;
        cudaEvent_t prev_gate = nullptr;
        const bool host_pack = std::getenv("SARLACC_HOST_PACK") != nullptr;     /* A/B: pack on the host as before */
        const bool pools_pinned = !reads->seq && pointer_is_pinned(reads->seq_pool) && pointer_is_pinned(reads->qual_pool);
        for (int64_t c0 = lo; c0 < hi;) {
            Slot& s = slots[which];
            drain(s, lay[which]);
            if (err.kind != ERR_NONE && err.at < c0) break;   /* found while draining: a serial run would have stopped there */
            /* size the chunk: the trace scratch budget may force fewer reads than `chunk` */
            int64_t c1 = std::min<int64_t>(hi, c0 + (c0 == lo ? first_chunk : chunk));
            s.h_lens.reserve(sizeof(int32_t) * (size_t)(c1 - c0));
            int maxlen = 0;
            scan_lengths(V, c0, c1, s.h_lens.as<int32_t>(), nthreads, err, maxlen);
            if (trace) {
                const long long fit = sub_chunk(P, maxlen, true, c1 - c0);
                if (fit < c1 - c0) {
                    c1 = c0 + fit;
                    maxlen = 0;
                    const int32_t* hl = s.h_lens.as<int32_t>();
                    for (int64_t i = 0; i < c1 - c0; ++i) maxlen = std::max(maxlen, (int)hl[i]);
                }
            }
            const long long m = c1 - c0;
            const int stride = std::max(8, (maxlen + 8) & ~7);   /* >= maxlen+1, multiple of 8 */
            if (host_pack) {
                s.h_rows.reserve(sizeof(uint16_t) * (size_t)m * stride);
                pack_rows(V, c0, c1, PT, s.h_lens.as<int32_t>(), stride, s.h_rows.as<uint16_t>(), nthreads, P.L > 0, err);
            }
            if (err.kind != ERR_NONE && err.at < c1) break;   /* a serial run would have stopped here */

            const OutLayout o = make_layout(mode, m, nref_scores, nsec, maxlen, P.L);
            lay[which] = o;
            s.d_rows.reserve(sizeof(uint16_t) * (size_t)m * stride);
            s.d_lens.reserve(sizeof(int32_t) * (size_t)m);
            s.d_out.reserve(o.total);
            s.h_out.reserve(o.total);
            /* uploads in chunk order (see the both-ends job below): chunk k's kernels should not wait for bytes of chunk k+1 */
            static const bool chain_uploads = std::getenv("SARLACC_NO_UPLOAD_CHAIN") == nullptr;
            if (prev_upload && chain_uploads) CUDA_CHECK(cudaStreamWaitEvent(s.st, prev_upload, 0));
            CUDA_CHECK(cudaMemcpyAsync(s.d_lens.p, s.h_lens.p, sizeof(int32_t) * (size_t)m, cudaMemcpyHostToDevice, s.st));
            PendingPack pending;
            if (host_pack) {
                CUDA_CHECK(cudaMemcpyAsync(s.d_rows.p, s.h_rows.p, sizeof(uint16_t) * (size_t)m * stride, cudaMemcpyHostToDevice, s.st));
            } else {
                stage_and_pack(V, c0, c1, PT, s.h_lens.as<int32_t>(), s.d_lens.as<int32_t>(), stride, s.d_rows.as<uint16_t>(), P.L > 0,
                               pools_pinned, s.raw, s.st, nthreads, &pending);
            }
            CUDA_CHECK(cudaEventRecord(s.t_h2d, s.st));
            prev_upload = s.t_h2d;
            launch_pending_pack(pending, s.st);
            uint8_t* d = s.d_out.as<uint8_t>();
            Outputs dev;
            dev.score = (mode == MODE_MULTI_GLOBAL && !want_all_scores) ? nullptr : reinterpret_cast<double*>(d + o.o_score);
            if (mode == MODE_TRACE_LOCAL) {
                dev.start = reinterpret_cast<int32_t*>(d + o.o_start);
                dev.end = reinterpret_cast<int32_t*>(d + o.o_end);
                dev.sec_start = reinterpret_cast<int32_t*>(d + o.o_ss);
                dev.sec_width = reinterpret_cast<int32_t*>(d + o.o_sw);
            }
            if (mode == MODE_MULTI_GLOBAL) {
                dev.best_id = reinterpret_cast<int32_t*>(d + o.o_bid);
                dev.best = reinterpret_cast<double*>(d + o.o_best);
                dev.next_best = reinterpret_cast<double*>(d + o.o_next);
            }
            if (mode == MODE_OPS_GLOBAL) {
                dev.nops = reinterpret_cast<int32_t*>(d + o.o_nops);
                dev.ops = d + o.o_ops;
                dev.ops_stride = o.ops_stride;
            }
            /* Barcode-length reads, score-only: walk the chunk in order of length (counting sort), so that the alignments a
             * warp works on side by side end together -- 24-row alignments otherwise spend half their steps waiting for
             * whichever lane finishes next (configs[3]: 0.11 s -> see profiles/r02_history.md).  Group g of the launch takes
             * positions g, g + NG, ... of the order: every second round of NG positions is turned round, so that a group's
             * reads add up to about the same number of rows (shortest + longest, ...) instead of the last groups getting
             * the longest read of every round. */
            const int32_t* d_order = nullptr;
            if (!trace && P.fast && maxlen < 48 && m > 1 && std::getenv("SARLACC_NO_LENGTH_ORDER") == nullptr) {
                s.h_order.reserve(sizeof(int32_t) * (size_t)m);
                s.d_order.reserve(sizeof(int32_t) * (size_t)m);
                const int32_t* hl = s.h_lens.as<int32_t>();
                int32_t* ho = s.h_order.as<int32_t>();
                std::vector<long long> start((size_t)maxlen + 2, 0);
                for (long long i = 0; i < m; ++i) ++start[(size_t)hl[i] + 1];
                for (int l = 0; l <= maxlen; ++l) start[(size_t)l + 1] += start[(size_t)l];
                for (long long i = 0; i < m; ++i) ho[start[(size_t)hl[i]]++] = (int32_t)i;
                const long long NG = launch_groups(P, maxlen, true, false);
                if (NG > 0 && std::getenv("SARLACC_NO_SERPENTINE") == nullptr) {
                    std::reverse(ho, ho + m);          /* longest first: a last, partial round holds the shortest reads */
                    for (long long r0 = NG; r0 < m; r0 += 2 * NG) std::reverse(ho + r0, ho + std::min(m, r0 + NG));
                }
                CUDA_CHECK(cudaMemcpyAsync(s.d_order.p, ho, sizeof(int32_t) * (size_t)m, cudaMemcpyHostToDevice, s.st));
                d_order = s.d_order.as<int32_t>();
            }
            /* chunks with tracebacks run their forward passes one after the other; score-only chunks may overlap, so that
             * the next chunk's first blocks fill the tail of this one's launch */
            static const bool gate_all = std::getenv("SARLACC_GATE_CHUNKS") != nullptr;
            if (prev_gate && (trace || gate_all)) CUDA_CHECK(cudaStreamWaitEvent(s.st, prev_gate, 0));
            run_device(P, D, s.scratch, s.st, s.d_rows.as<uint16_t>(), s.d_lens.as<int32_t>(), m, stride, maxlen, trace, dev, sms,
                       nullptr, nullptr, nullptr, nullptr, d_order);
            CUDA_CHECK(cudaEventRecord(s.gate, s.st));
            prev_gate = s.gate;
            CUDA_CHECK(cudaMemcpyAsync(s.h_out.p, s.d_out.p, o.total, cudaMemcpyDeviceToHost, s.st));
            CUDA_CHECK(cudaEventRecord(s.done, s.st));
            s.n = m;
            s.lo = c0;
            s.busy = true;
            which = (which + 1) % kSlots;
            c0 = c1;
        }
        for (int k = 0; k < kSlots; ++k) drain(slots[(which + k) % kSlots], lay[(which + k) % kSlots]);   /* oldest first */
    }
};

std::vector<int> configured_devices() {
    std::lock_guard<std::mutex> lock(g_cfg_mutex);
    if (g_devices.empty()) return std::vector<int>{0};
    return g_devices;
}

int require_device() {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        cudaGetLastError();
        return fail(std::string("sarlacc_b200 requires a CUDA device (no CPU fallback exists): ") +
                    (e != cudaSuccess ? cudaGetErrorString(e) : "no device found"));
    }
    return 0;
}

/* Degenerate references (L == 0) never reach a kernel: src/reference_align.cpp:82-90 runs no column, the score
 * is scores[len] of column 0 and querymap returns (0,0) (:308-310). */
double col0_host(bool local, double gop, double ge, int64_t i) {
    if (local || i == 0) return 0.0;
    return -gop - ge * (double)(i - 1);
}

/* Unrecognized reference bases (src/reference_align.cpp:211) surface at the first read that has any base:
 * column by column, so a bad quality in that read wins unless the very first column is the bad one
 * (:184-217).  With several references the unfused R loop runs one .Call per barcode, so read errors (raised
 * in the first call) win over a bad later barcode. */
void apply_reference_errors(FirstError& err, const ReadView& V, int64_t n, const Plan& P, int nref) {
    int bfirst = -1;
    for (int b = 0; b < nref; ++b) {
        if (P.bad_col[b] >= 0) { bfirst = b; break; }
    }
    if (bfirst < 0) return;
    int64_t i0 = -1;
    for (int64_t i = 0; i < n; ++i) {
        if (V.seq_len(i) != V.qual_len(i)) break;
        if (V.seq_len(i) > 0) { i0 = i; break; }
    }
    if (i0 < 0) return;
    if (bfirst == 0) {
        if (err.kind == ERR_NONE || err.at > i0) {
            err.at = i0;
            err.kind = ERR_REF;
        } else if (err.at == i0 && P.bad_col[0] == 0) {
            err.kind = ERR_REF;
        }
    } else if (err.kind == ERR_NONE) {
        err.at = i0;
        err.kind = ERR_REF;
    }
}

/* Shared driver of all host-buffer entry points. */
int run_host(const sarlacc_reads* reads, const sarlacc_encoding* encoding, double go, double ge,
        const char* const* refs, int nref, Mode mode, int nsec, const int32_t* sec_starts, const int32_t* sec_ends,
        HostOutputs out, bool want_all_scores)
{
    if (!reads) return fail("reads must not be NULL");
    Encoding enc;
    const char* msg = build_encoding(encoding, enc);
    if (msg) return fail(msg);   /* reference_align's constructor runs before any read is touched */
    const int64_t n = reads->n;
    const bool local = (mode == MODE_SCORE_LOCAL || mode == MODE_TRACE_LOCAL);
    int L = 0;
    if (nref > 0) {
        L = (int)std::strlen(refs[0]);
        for (int b = 1; b < nref; ++b) {
            if ((int)std::strlen(refs[b]) != L) return fail("all barcodes must have the same length for a fused pass");
        }
    }
    Plan P;
    build_plan(P, enc, refs, nref, L, local, go, ge, mode == MODE_TRACE_LOCAL || mode == MODE_OPS_GLOBAL);
    if (nsec > 0) {
        P.sec_starts.assign(sec_starts, sec_starts + nsec);
        P.sec_ends.assign(sec_ends, sec_ends + nsec);
        for (int s = 0; s < nsec; ++s) {
            /* the reference indexes its mapping deque unchecked (:326-350); refuse what would be out of bounds there */
            if (P.sec_starts[s] < 0 || P.sec_starts[s] > L || P.sec_ends[s] < 0 || P.sec_ends[s] > L) {
                return fail("section bounds outside the adaptor");
            }
        }
    }
    if (n == 0) return 0;
    ReadView V{reads};

    if (L == 0 || nref == 0) {
        /* no DP column exists; only the length check of the entry loop can fail */
        for (int64_t i = 0; i < n; ++i) {
            if (V.seq_len(i) != V.qual_len(i)) return fail(err_text(ERR_LEN));
            const double s = col0_host(local, P.gop, P.ge, V.seq_len(i));
            if (mode == MODE_MULTI_GLOBAL) {
                out.best_id[i] = 0;
                out.best[i] = -std::numeric_limits<double>::infinity();
                out.next_best[i] = -std::numeric_limits<double>::infinity();
            } else {
                out.score[i] = s;
            }
            if (mode == MODE_TRACE_LOCAL) {
                out.start[i] = 0;
                out.end[i] = 0;
                for (int sct = 0; sct < nsec; ++sct) {
                    out.sec_start[(size_t)sct * n + i] = 1;
                    out.sec_width[(size_t)sct * n + i] = 0;
                }
            }
            if (mode == MODE_OPS_GLOBAL) (*out.ops)[i].assign((size_t)V.seq_len(i), (uint8_t)'I');
        }
        return 0;
    }

    if (require_device()) return 1;
    std::vector<int> devs = configured_devices();
    if ((int64_t)devs.size() > n) devs.resize((size_t)std::max<int64_t>(1, n));
    const int nd = (int)devs.size();
    std::vector<DeviceJob> jobs(nd);
    for (int d = 0; d < nd; ++d) {
        DeviceJob& J = jobs[d];
        J.device = devs[d];
        J.lo = n * d / nd;
        J.hi = n * (d + 1) / nd;
        J.reads = reads;
        J.plan = &P;
        J.mode = mode;
        J.n_total = n;
        J.out = out;
        J.want_all_scores = want_all_scores;
        J.nthreads = host_threads_for(nd);
    }
    if (nd == 1) {
        jobs[0].run();
    } else {
        std::vector<std::thread> pool;
        for (int d = 0; d < nd; ++d) pool.emplace_back([&jobs, d] { jobs[d].run(); });
        for (auto& t : pool) t.join();
    }
    for (int d = 0; d < nd; ++d) {
        if (!jobs[d].cuda_error.empty()) return fail(jobs[d].cuda_error);
    }
    FirstError err;
    for (int d = 0; d < nd; ++d) err.merge(jobs[d].err);

    apply_reference_errors(err, V, n, P, nref);
    if (err.kind != ERR_NONE) return fail(err_text(err.kind));
    return 0;
}

/* ---- fused both-ends entry ----------------------------------------------------------------------
 * What .align_AA_internal (R/adaptorAlign.R:178-199) does with four .Calls -- (adaptor1, front), (adaptor2, back),
 * (adaptor1, back), (adaptor2, front) -- plus .resolve_strand, the row selection and adaptorAlign's adaptor2
 * coordinate flip (:66-71), in one pass: both window sets are packed and uploaded once, the four alignments run
 * back to back on the device, and only the selected rows come back. */
struct PairOutputs {
    uint8_t* reversed;
    double* score[2];
    int32_t* start[2];
    int32_t* end[2];
    int32_t* sec_start[2];
    int32_t* sec_width[2];
};

struct FinalLayout {
    size_t o_rev, o_score[2], o_start[2], o_end[2], o_ss[2], o_sw[2], total;
};

/* Host-side phases of the last fused both-ends call on this thread's job 0 (bench.py reports them per rank):
 * staging / length scans, enqueueing, waiting for results + copying them out, total; milliseconds. */
std::mutex g_pair_timing_mutex;
double g_pair_timing[6] = {0, 0, 0, 0, 0, 0};
long long g_pair_upload_bytes = 0;

struct PairJob {
    int device = 0;
    int64_t lo = 0, hi = 0, n_total = 0;
    const sarlacc_reads* front = nullptr;
    const sarlacc_reads* back = nullptr;
    const Plan* plan[2] = {nullptr, nullptr};   /* adaptor1, adaptor2 */
    const int32_t* width = nullptr;
    int tolerance = 0;          /* > 0: `front` holds WHOLE reads; both windows are cut (and the back one reverse-complemented) by the packer */
    int32_t* width_out = nullptr;
    PairOutputs out;
    int nthreads = 1;
    FirstError err_front, err_back;
    std::string cuda_error;

    void run() {
        try {
            run_inner();
        } catch (CudaError& e) {
            cuda_error = e.msg;
        } catch (std::exception& e) {
            cuda_error = std::string("internal error: ") + e.what();
        }
    }

    void run_inner() {
        CUDA_CHECK(cudaSetDevice(device));
        const int sms = device_sm_count(device);
        ReadView VF{front}, VB{back};
        if (tolerance > 0) {
            VF.tol = tolerance;
            VB = ReadView{front, tolerance, true};
        }
        PackTables PF, PB;
        build_pack_tables(PF, front->seq_encoding, *plan[0]->enc);
        build_pack_tables(PB, VB.R->seq_encoding, *plan[0]->enc);
        const int nsec[2] = {(int)plan[0]->sec_starts.size(), (int)plan[1]->sec_starts.size()};
        DeviceCache& cache = DeviceCache::acquire(device);
        struct Release {
            DeviceCache& c;
            ~Release() { c.release(); }
        } release{cache};
        Slot* slots = cache.slots;
        DevPlan D[2];
        struct FreePlans {
            DevPlan* d;
            ~FreePlans() { d[0].buf.release(); d[1].buf.release(); }
        } free_plans{D};
        D[0].upload(*plan[0], slots[0].st);
        D[1].upload(*plan[1], slots[0].st);

        /* chunk = a whole number of grid-fulls for both adaptors' kernels (no tail round); the first chunk is a quarter
         * of that so that the device starts early */
        const long long g_a1 = plan_groups(*plan[0], true), g_a2 = plan_groups(*plan[1], true);
        long long chunk = chunk_for(1 << 17, g_a1, g_a2);
        long long first_chunk = (hi - lo > chunk) ? std::max<long long>(1, whole_rounds(chunk / 4, std::max(g_a1, g_a2) > 0 ? (g_a1 > 0 ? g_a1 : g_a2) : 0)) : chunk;
        const char* ce = std::getenv("SARLACC_CHUNK");
        if (ce && std::atoll(ce) > 0) first_chunk = chunk = std::atoll(ce);
        auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
        FinalLayout lay[kMaxSlots];
        int which = 0;

        const bool dbg = std::getenv("SARLACC_DEBUG_TIMING") != nullptr;
        auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        const double t_start = now();
        double t_h2d_ms = 0, t_dev_ms = 0;      /* summed over chunks: upload (incl. device packer) / kernels + copy back, from CUDA events */
        size_t bytes_sent = 0;
        auto drain = [&](Slot& s, const FinalLayout& o) {
            if (!s.busy) return;
            Range nvtx("sarlacc: wait for chunk + copy out");
            CUDA_CHECK(cudaEventSynchronize(s.done));
            {
                const long long fb = s.raw.first_bad(), fb2 = s.raw2.first_bad();
                if (fb >= 0) err_front.offer(s.lo + fb, ERR_QUAL);
                if (fb2 >= 0) err_back.offer(s.lo + fb2, ERR_QUAL);
            }
            {
                float a = 0, b = 0;
                cudaEventElapsedTime(&a, s.t_begin, s.t_h2d);
                cudaEventElapsedTime(&b, s.t_h2d, s.t_end);
                t_h2d_ms += a;
                t_dev_ms += b;
                bytes_sent += s.up.sent;
                const double t_now = now();
                if (dbg) std::fprintf(stderr, "[sarlacc]   chunk at %lld (%lld reads): H2D %.2f ms (%.1f MB%s), kernels + D2H %.2f ms, host clock %.1f ms\n",
                                      (long long)s.lo, s.n, a, s.up.sent / 1e6, s.up.sent < s.up.raw ? ", bases as 4-bit codes" : "", b, (t_now - t_start) * 1e3);
            }
            const uint8_t* h = s.h_out.as<uint8_t>();
            const long long m = s.n;
            const int64_t g0 = s.lo;
            std::memcpy(out.reversed + g0, h + o.o_rev, (size_t)m);
            for (int k = 0; k < 2; ++k) {
                std::memcpy(out.score[k] + g0, h + o.o_score[k], sizeof(double) * m);
                std::memcpy(out.start[k] + g0, h + o.o_start[k], sizeof(int32_t) * m);
                std::memcpy(out.end[k] + g0, h + o.o_end[k], sizeof(int32_t) * m);
                for (int sct = 0; sct < nsec[k]; ++sct) {
                    std::memcpy(out.sec_start[k] + (size_t)sct * n_total + g0, h + o.o_ss[k] + sizeof(int32_t) * (size_t)sct * m, sizeof(int32_t) * m);
                    std::memcpy(out.sec_width[k] + (size_t)sct * n_total + g0, h + o.o_sw[k] + sizeof(int32_t) * (size_t)sct * m, sizeof(int32_t) * m);
                }
            }
            s.busy = false;
        };

        double t_drain = 0, t_pack = 0, t_enq = 0;
        cudaEvent_t prev_gate = nullptr;
        cudaEvent_t prev_upload = nullptr;
        cudaEvent_t pair_free[2] = {nullptr, nullptr};     /* behind the tracebacks of the last chunk that used record group p */
        int pair_turn = 0;
        const bool host_pack = std::getenv("SARLACC_HOST_PACK") != nullptr;     /* A/B: pack on the host as before */
        /* SARLACC_PACK_SEQ=1: send the bases as 4-bit codes (see stage_host).  Off by default: on the 8-GPU box, where the
         * ranks do wait for the link, the 4 host cores per rank need longer for the pass over the sequence bytes than the
         * link saves (61 -> 91 ms per step, profiles/r02_history.md); a host with cores to spare is the use case. */
        const char* nib_env = std::getenv("SARLACC_PACK_SEQ");
        const bool nib = nib_env && std::atoi(nib_env) != 0;
        const bool pinned_f = !VF.R->seq && pointer_is_pinned(VF.R->seq_pool) && pointer_is_pinned(VF.R->qual_pool);
        const bool pinned_b = !VB.R->seq && pointer_is_pinned(VB.R->seq_pool) && pointer_is_pinned(VB.R->qual_pool);
        for (int64_t c0 = lo; c0 < hi;) {
            Slot& s = slots[which];
            double t0 = now();
            drain(s, lay[which]);
            t_drain += now() - t0;
            if ((err_front.kind != ERR_NONE && err_front.at < c0) || (err_back.kind != ERR_NONE && err_back.at < c0)) break;
            t0 = now();
            /* chunk sizes: a quarter, a half, then whole chunks -- each upload is hidden behind the chunk before it (a whole
             * chunk right after the quarter left the device idle for half an upload) -- and the tail of the range split so
             * that the last chunk, whose copy back nothing hides, is a small one */
            int64_t step = c0 == lo ? first_chunk : (c0 == lo + first_chunk && first_chunk < chunk ? std::max<int64_t>(first_chunk, chunk / 2) : chunk);
            if (first_chunk < chunk && hi - c0 > step / 4 && hi - c0 <= step + step / 4) step = hi - c0 - step / 4;
            int64_t c1 = std::min<int64_t>(hi, c0 + step);
            s.h_lens.reserve(sizeof(int32_t) * (size_t)(c1 - c0));
            s.h_lens2.reserve(sizeof(int32_t) * (size_t)(c1 - c0));
            int maxf = 0, maxb = 0;
            scan_lengths(VF, c0, c1, s.h_lens.as<int32_t>(), nthreads, err_front, maxf);
            scan_lengths(VB, c0, c1, s.h_lens2.as<int32_t>(), nthreads, err_back, maxb);
            {
                const int mx = std::max(maxf, maxb);
                const long long fit = std::max<long long>(1, (long long)(2 * scratch_budget_bytes() / pair_scratch_per_read(*plan[0], *plan[1], mx)));
                if (fit < c1 - c0) {
                    c1 = c0 + fit;
                    maxf = maxb = 0;
                    for (int64_t i = 0; i < c1 - c0; ++i) {
                        maxf = std::max(maxf, (int)s.h_lens.as<int32_t>()[i]);
                        maxb = std::max(maxb, (int)s.h_lens2.as<int32_t>()[i]);
                    }
                }
            }
            const long long m = c1 - c0;
            const int stride_f = std::max(8, (maxf + 8) & ~7), stride_b = std::max(8, (maxb + 8) & ~7);
            if (host_pack) {
                s.h_rows.reserve(sizeof(uint16_t) * (size_t)m * stride_f);
                s.h_rows2.reserve(sizeof(uint16_t) * (size_t)m * stride_b);
                pack_rows(VF, c0, c1, PF, s.h_lens.as<int32_t>(), stride_f, s.h_rows.as<uint16_t>(), nthreads, true, err_front);
                pack_rows(VB, c0, c1, PB, s.h_lens2.as<int32_t>(), stride_b, s.h_rows2.as<uint16_t>(), nthreads, true, err_back);
            }
            if ((err_front.kind != ERR_NONE && err_front.at < c1) || (err_back.kind != ERR_NONE && err_back.at < c1)) break;
            t_pack += now() - t0;
            t0 = now();

            /* device buffers */
            s.d_rows.reserve(sizeof(uint16_t) * (size_t)m * stride_f);
            s.d_rows2.reserve(sizeof(uint16_t) * (size_t)m * stride_b);
            s.d_lens.reserve(sizeof(int32_t) * (size_t)m);
            s.d_lens2.reserve(sizeof(int32_t) * (size_t)m);
            size_t at = 0;
            FinalLayout F;
            at = 0;
            F.o_rev = at; at += al((size_t)m);
            for (int k = 0; k < 2; ++k) {
                F.o_score[k] = at; at += al(sizeof(double) * m);
                F.o_start[k] = at; at += al(sizeof(int32_t) * m);
                F.o_end[k] = at; at += al(sizeof(int32_t) * m);
                F.o_ss[k] = at; at += al(sizeof(int32_t) * m * std::max(1, nsec[k]));
                F.o_sw[k] = at; at += al(sizeof(int32_t) * m * std::max(1, nsec[k]));
            }
            F.total = at;
            lay[which] = F;
            s.d_tmp.reserve(sizeof(double) * 4 * (size_t)m);      /* the four forward passes' scores */
            s.d_out.reserve(F.total);
            s.h_out.reserve(F.total);
            /* host work first (offsets, staging or 4-bit coding of the bases), so that the stream holds nothing but copies
             * between t_begin and t_h2d */
            Staged gf, gb;
            if (!host_pack) {
                gf = stage_host(VF, c0, c1, PF, s.h_lens.as<int32_t>(), pinned_f, s.raw, nthreads, nib);
                gb = stage_host(VB, c0, c1, PB, s.h_lens2.as<int32_t>(), pinned_b, s.raw2, nthreads, nib);
            }
            /* uploads in chunk order: copies of different streams share the link, and chunk k's forward passes should not
             * wait for bytes of chunk k+1 (at the start of a call three chunks are enqueued at once) */
            static const bool chain_uploads = std::getenv("SARLACC_NO_UPLOAD_CHAIN") == nullptr;
            if (prev_upload && chain_uploads) CUDA_CHECK(cudaStreamWaitEvent(s.st, prev_upload, 0));
            CUDA_CHECK(cudaEventRecord(s.t_begin, s.st));
            CUDA_CHECK(cudaMemcpyAsync(s.d_lens.p, s.h_lens.p, sizeof(int32_t) * (size_t)m, cudaMemcpyHostToDevice, s.st));
            CUDA_CHECK(cudaMemcpyAsync(s.d_lens2.p, s.h_lens2.p, sizeof(int32_t) * (size_t)m, cudaMemcpyHostToDevice, s.st));
            PendingPack pf, pb;      /* both uploads first, then both packers (see stage_and_pack) */
            s.up = UploadBytes();
            if (host_pack) {
                CUDA_CHECK(cudaMemcpyAsync(s.d_rows.p, s.h_rows.p, sizeof(uint16_t) * (size_t)m * stride_f, cudaMemcpyHostToDevice, s.st));
                CUDA_CHECK(cudaMemcpyAsync(s.d_rows2.p, s.h_rows2.p, sizeof(uint16_t) * (size_t)m * stride_b, cudaMemcpyHostToDevice, s.st));
            } else {
                stage_enqueue(gf, VF, PF, s.d_lens.as<int32_t>(), stride_f, s.d_rows.as<uint16_t>(), true, s.raw, s.st, &pf, &s.up);
                stage_enqueue(gb, VB, PB, s.d_lens2.as<int32_t>(), stride_b, s.d_rows2.as<uint16_t>(), true, s.raw2, s.st, &pb, &s.up);
            }
            if (width || tolerance > 0) {
                s.h_width.reserve(sizeof(int32_t) * (size_t)m);
                s.d_width.reserve(sizeof(int32_t) * (size_t)m);
                if (tolerance > 0) {
                    int32_t* hw = s.h_width.as<int32_t>();
                    for (int64_t i = 0; i < m; ++i) {
                        hw[i] = (int32_t)VF.full_seq_len(c0 + i);
                        if (width_out) width_out[c0 + i] = hw[i];
                    }
                } else {
                    std::memcpy(s.h_width.p, width + c0, sizeof(int32_t) * (size_t)m);
                }
                CUDA_CHECK(cudaMemcpyAsync(s.d_width.p, s.h_width.p, sizeof(int32_t) * (size_t)m, cudaMemcpyHostToDevice, s.st));
            }
            CUDA_CHECK(cudaEventRecord(s.t_h2d, s.st));      /* uploads done; the device packers follow */
            prev_upload = s.t_h2d;
            launch_pending_pack(pf, s.st);
            launch_pending_pack(pb, s.st);
            /* consecutive chunks use different record sets (pair_free below), so their forward passes need not wait for
             * each other: the next chunk's first blocks start in the tail of this chunk's last launches, as the sub-ranges
             * of a device-resident chunk do (SARLACC_GATE_CHUNKS=1 keeps them in order, for comparison) */
            static const bool gate_all = std::getenv("SARLACC_GATE_CHUNKS") != nullptr;
            if (prev_gate && gate_all) CUDA_CHECK(cudaStreamWaitEvent(s.st, prev_gate, 0));
            uint8_t* d = s.d_out.as<uint8_t>();
            PairDeviceOut po;
            po.reversed = d + F.o_rev;
            for (int k = 0; k < 2; ++k) {
                po.score[k] = reinterpret_cast<double*>(d + F.o_score[k]);
                po.start[k] = reinterpret_cast<int32_t*>(d + F.o_start[k]);
                po.end[k] = reinterpret_cast<int32_t*>(d + F.o_end[k]);
                po.sec_start[k] = reinterpret_cast<int32_t*>(d + F.o_ss[k]);
                po.sec_width[k] = reinterpret_cast<int32_t*>(d + F.o_sw[k]);
            }
            po.pitch = m;
            const DevPlan* Dp[2] = {&D[0], &D[1]};
            const int pg = pair_turn;
            pair_turn ^= 1;
            if (pair_free[pg]) CUDA_CHECK(cudaStreamWaitEvent(s.st, pair_free[pg], 0));    /* chunk k-2's tracebacks read these records */
            pair_free[pg] = s.tb_ev[0];
            run_pair_device(plan, Dp, cache.pair[pg], s.st, s.tb, s.fwd_ev[0], s.tb_ev[0],
                            s.d_rows.as<uint16_t>(), s.d_lens.as<int32_t>(), stride_f, s.d_rows2.as<uint16_t>(), s.d_lens2.as<int32_t>(), stride_b,
                            m, std::max(maxf, maxb), (width || tolerance > 0) ? s.d_width.as<int32_t>() : nullptr,
                            s.d_tmp.as<double>(), po, sms);
            CUDA_CHECK(cudaEventRecord(s.gate, s.st));     /* behind the fourth forward pass; the tracebacks run beside the next chunk */
            prev_gate = s.gate;
            CUDA_CHECK(cudaStreamWaitEvent(s.st, s.tb_ev[0], 0));
            CUDA_CHECK(cudaMemcpyAsync(s.h_out.p, s.d_out.p, F.total, cudaMemcpyDeviceToHost, s.st));
            CUDA_CHECK(cudaEventRecord(s.t_end, s.st));
            CUDA_CHECK(cudaEventRecord(s.done, s.st));
            s.n = m;
            s.lo = c0;
            s.busy = true;
            which = (which + 1) % kSlots;
            c0 = c1;
            t_enq += now() - t0;
        }
        double t0 = now();
        for (int k = 0; k < kSlots; ++k) drain(slots[(which + k) % kSlots], lay[(which + k) % kSlots]);   /* oldest first */
        t_drain += now() - t0;
        if (dbg) std::fprintf(stderr, "[sarlacc] pair job dev %d: pack %.1f ms, enqueue %.1f ms, wait+copy-out %.1f ms\n", device, t_pack * 1e3, t_enq * 1e3, t_drain * 1e3);
        {
            std::lock_guard<std::mutex> lock(g_pair_timing_mutex);
            g_pair_timing[0] = t_pack * 1e3;
            g_pair_timing[1] = t_enq * 1e3;
            g_pair_timing[2] = t_drain * 1e3;
            g_pair_timing[3] = (now() - t_start) * 1e3;
            g_pair_timing[4] = t_h2d_ms;
            g_pair_timing[5] = t_dev_ms;
            g_pair_upload_bytes = (long long)bytes_sent;
        }
    }
};

}  // namespace

/* =================================================================================================
 * C ABI
 * ================================================================================================= */

extern "C" {

const char* sarlacc_last_error(void) { return g_error.c_str(); }

const char* sarlacc_version(void) { return "sarlacc_b200 0.1 (sm_100a)"; }

int sarlacc_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return count;
}

int sarlacc_set_devices(const int* devices, int ndevices) {
    if (ndevices < 0 || (ndevices > 0 && !devices)) return fail("invalid device list");
    const int count = sarlacc_device_count();
    for (int i = 0; i < ndevices; ++i) {
        if (devices[i] < 0 || devices[i] >= count) return fail("device index out of range");
    }
    std::lock_guard<std::mutex> lock(g_cfg_mutex);
    g_devices.assign(devices, devices + ndevices);
    return 0;
}

int sarlacc_set_host_threads(int nthreads) {
    if (nthreads < 0) return fail("nthreads must be >= 0");
    g_host_threads = nthreads;
    return 0;
}

void sarlacc_trim_device_memory(void) {
    DevPool::instance().trim();
    sarlacc::threshold_trim();
}

void sarlacc_last_pair_timing(double* ms6) {
    if (!ms6) return;
    std::lock_guard<std::mutex> lock(g_pair_timing_mutex);
    for (int k = 0; k < 6; ++k) ms6[k] = g_pair_timing[k];
}

int64_t sarlacc_last_pair_upload_bytes(void) {
    std::lock_guard<std::mutex> lock(g_pair_timing_mutex);
    return g_pair_upload_bytes;
}

int64_t sarlacc_kernel_launches(int reset) {
    const long long v = g_launches.load();
    if (reset) g_launches.store(0);
    return v;
}

int sarlacc_adaptor_align(const sarlacc_reads* reads, const sarlacc_encoding* encoding,
        double gapopen, double gapext, const char* adaptor,
        int nsec, const int32_t* sec_starts, const int32_t* sec_ends,
        double* score, int32_t* start, int32_t* end, int32_t* sec_start, int32_t* sec_width)
{
    if (!adaptor) return fail("adaptor sequence should be a string");
    if (nsec < 0) return fail("section starts and ends should have the same length");
    HostOutputs out;
    out.score = score;
    out.start = start;
    out.end = end;
    out.sec_start = sec_start;
    out.sec_width = sec_width;
    const char* refs[1] = {adaptor};
    return run_host(reads, encoding, gapopen, gapext, refs, 1, MODE_TRACE_LOCAL, nsec, sec_starts, sec_ends, out, false);
}

int sarlacc_adaptor_align_score_only(const sarlacc_reads* reads, const sarlacc_encoding* encoding,
        double gapopen, double gapext, const char* adaptor, double* score)
{
    if (!adaptor) return fail("adaptor sequence should be a string");
    HostOutputs out;
    out.score = score;
    const char* refs[1] = {adaptor};
    return run_host(reads, encoding, gapopen, gapext, refs, 1, MODE_SCORE_LOCAL, 0, nullptr, nullptr, out, false);
}

int sarlacc_barcode_align(const sarlacc_reads* reads, const sarlacc_encoding* encoding,
        double gapopen, double gapext, const char* reference, double* score)
{
    if (!reference) return fail("barcode sequence should be a string");
    HostOutputs out;
    out.score = score;
    const char* refs[1] = {reference};
    return run_host(reads, encoding, gapopen, gapext, refs, 1, MODE_SCORE_GLOBAL, 0, nullptr, nullptr, out, false);
}

int sarlacc_barcode_align_multi(const sarlacc_reads* reads, const sarlacc_encoding* encoding,
        double gapopen, double gapext, const char* const* barcodes, int nbarcodes,
        int32_t* best_id, double* best, double* next_best, double* all_scores)
{
    if (nbarcodes < 0 || (nbarcodes > 0 && !barcodes)) return fail("barcode sequence should be a string");
    HostOutputs out;
    out.score = all_scores;
    out.best_id = best_id;
    out.best = best;
    out.next_best = next_best;
    return run_host(reads, encoding, gapopen, gapext, barcodes, nbarcodes, MODE_MULTI_GLOBAL, 0, nullptr, nullptr, out, all_scores != nullptr);
}

int sarlacc_general_align(const sarlacc_reads* reads, const sarlacc_encoding* encoding,
        double gapopen, double gapext, const char* reference, int edit_only,
        double* score, int32_t* edit, char* ref_aln, char* query_aln, int64_t aln_stride)
{
    if (!reference) return fail("reference sequence should be a string");
    if (!reads) return fail("reads must not be NULL");
    const int64_t n = reads->n;
    std::vector<std::vector<uint8_t> > ops((size_t)std::max<int64_t>(n, 0));
    HostOutputs out;
    out.score = score;
    out.ops = &ops;
    const char* refs[1] = {reference};
    int rc = run_host(reads, encoding, gapopen, gapext, refs, 1, MODE_OPS_GLOBAL, 0, nullptr, nullptr, out, false);
    if (rc) return rc;
    /* fill_strings + the edit-distance loop of src/general_align.cpp:44-57, from the device's operation
     * list and the ORIGINAL characters (the packed rows do not keep non-ACGT read characters). */
    ReadView V{reads};
    const size_t L = std::strlen(reference);
    static const char decode[16] = {'-', 'A', 'C', 'M', 'G', 'R', 'S', 'V', 'T', 'W', 'Y', 'H', 'K', 'D', 'B', 'N'};
    for (int64_t i = 0; i < n; ++i) {
        const std::vector<uint8_t>& op = ops[(size_t)i];
        const uint8_t* s = V.seq(i);
        const size_t nop = op.size();
        if (!edit_only && (int64_t)nop + 1 > aln_stride) return fail("aln_stride too small for the alignment strings");
        size_t ri = 0, qi = 0;
        int32_t ed = 0;
        for (size_t x = 0; x < nop; ++x) {
            const uint8_t o = op[nop - 1 - x];   /* device wrote back to front */
            char rc_ = '-', qc = '-';
            if (o != 'I') rc_ = reference[ri++];
            if (o != 'D') {
                const uint8_t raw = s[qi++];
                qc = (reads->seq_encoding == SARLACC_SEQ_BIOSTRINGS) ? (raw < 16 ? decode[raw] : (raw == 16 ? '-' : (raw == 32 ? '+' : '.'))) : (char)raw;
            }
            if (rc_ != qc) ++ed;
            if (!edit_only) {
                ref_aln[i * aln_stride + (int64_t)x] = rc_;
                query_aln[i * aln_stride + (int64_t)x] = qc;
            }
        }
        (void)L;
        edit[i] = ed;
        if (!edit_only) {
            ref_aln[i * aln_stride + (int64_t)nop] = '\0';
            query_aln[i * aln_stride + (int64_t)nop] = '\0';
        }
    }
    return 0;
}

static int align_pair(const sarlacc_reads* front, const sarlacc_reads* back, int tolerance, const sarlacc_encoding* encoding,
        double gapopen, double gapext, const char* adaptor1, const char* adaptor2,
        int nsec1, const int32_t* sec_starts1, const int32_t* sec_ends1,
        int nsec2, const int32_t* sec_starts2, const int32_t* sec_ends2,
        const int32_t* read_width, int32_t* width_out, uint8_t* reversed,
        double* score1, int32_t* start1, int32_t* end1, int32_t* sec_start1, int32_t* sec_width1,
        double* score2, int32_t* start2, int32_t* end2, int32_t* sec_start2, int32_t* sec_width2)
{
    if (!front || !back) return fail("reads must not be NULL");
    if (!adaptor1 || !adaptor2) return fail("adaptor sequence should be a string");
    if (front->n != back->n) return fail("front and back windows should have the same length");
    if (nsec1 < 0 || nsec2 < 0) return fail("section starts and ends should have the same length");
    Encoding enc;
    const char* msg = build_encoding(encoding, enc);
    if (msg) return fail(msg);
    const int64_t n = front->n;
    const int L1 = (int)std::strlen(adaptor1), L2 = (int)std::strlen(adaptor2);
    if (n == 0) return 0;
    if ((L1 == 0 || L2 == 0) && tolerance > 0) return fail("the fused whole-read entry needs two non-empty adaptors");
    if (L1 == 0 || L2 == 0) {
        /* degenerate adaptors: compose the four reference calls on the host side of the ABI */
        std::vector<double> sc[4];
        std::vector<int32_t> st[4], en[4], ss[4], sw[4];
        const sarlacc_reads* rd[4] = {front, back, back, front};
        const char* ad[4] = {adaptor1, adaptor2, adaptor1, adaptor2};
        const int ns[4] = {nsec1, nsec2, nsec1, nsec2};
        const int32_t* s0[4] = {sec_starts1, sec_starts2, sec_starts1, sec_starts2};
        const int32_t* e0[4] = {sec_ends1, sec_ends2, sec_ends1, sec_ends2};
        for (int r = 0; r < 4; ++r) {
            sc[r].resize(n); st[r].resize(n); en[r].resize(n);
            ss[r].resize((size_t)std::max(1, ns[r]) * n); sw[r].resize((size_t)std::max(1, ns[r]) * n);
            int rc = sarlacc_adaptor_align(rd[r], encoding, gapopen, gapext, ad[r], ns[r], s0[r], e0[r],
                                           sc[r].data(), st[r].data(), en[r].data(), ss[r].data(), sw[r].data());
            if (rc) return rc;
        }
        for (int64_t i = 0; i < n; ++i) {
            const double f = std::max(sc[0][i], 0.0) + std::max(sc[1][i], 0.0);
            const double r = std::max(sc[2][i], 0.0) + std::max(sc[3][i], 0.0);
            const bool rev = f < r;
            reversed[i] = rev ? 1 : 0;
            const int a = rev ? 2 : 0, b = rev ? 3 : 1;
            score1[i] = sc[a][i]; start1[i] = st[a][i]; end1[i] = en[a][i];
            for (int s = 0; s < nsec1; ++s) { sec_start1[(size_t)s * n + i] = ss[a][(size_t)s * n + i]; sec_width1[(size_t)s * n + i] = sw[a][(size_t)s * n + i]; }
            score2[i] = sc[b][i];
            int x = st[b][i], y = en[b][i];
            if (read_width) { x = read_width[i] - x + 1; y = read_width[i] - y + 1; }
            start2[i] = x; end2[i] = y;
            for (int s = 0; s < nsec2; ++s) { sec_start2[(size_t)s * n + i] = ss[b][(size_t)s * n + i]; sec_width2[(size_t)s * n + i] = sw[b][(size_t)s * n + i]; }
        }
        return 0;
    }
    Plan P[2];
    const char* r1[1] = {adaptor1};
    const char* r2[1] = {adaptor2};
    build_plan(P[0], enc, r1, 1, L1, true, gapopen, gapext);
    build_plan(P[1], enc, r2, 1, L2, true, gapopen, gapext);
    const int nsecs[2] = {nsec1, nsec2};
    const int32_t* sst[2] = {sec_starts1, sec_starts2};
    const int32_t* sen[2] = {sec_ends1, sec_ends2};
    const int Ls[2] = {L1, L2};
    for (int k = 0; k < 2; ++k) {
        if (nsecs[k] > 0) {
            P[k].sec_starts.assign(sst[k], sst[k] + nsecs[k]);
            P[k].sec_ends.assign(sen[k], sen[k] + nsecs[k]);
            for (int s = 0; s < nsecs[k]; ++s) {
                if (sst[k][s] < 0 || sst[k][s] > Ls[k] || sen[k][s] < 0 || sen[k][s] > Ls[k]) return fail("section bounds outside the adaptor");
            }
        }
    }
    if (require_device()) return 1;
    std::vector<int> devs = configured_devices();
    if ((int64_t)devs.size() > n) devs.resize((size_t)std::max<int64_t>(1, n));
    const int nd = (int)devs.size();
    std::vector<PairJob> jobs(nd);
    for (int d = 0; d < nd; ++d) {
        PairJob& J = jobs[d];
        J.device = devs[d];
        J.lo = n * d / nd;
        J.hi = n * (d + 1) / nd;
        J.n_total = n;
        J.front = front;
        J.back = back;
        J.plan[0] = &P[0];
        J.plan[1] = &P[1];
        J.width = read_width;
        J.tolerance = tolerance;
        J.width_out = width_out;
        J.out.reversed = reversed;
        J.out.score[0] = score1; J.out.start[0] = start1; J.out.end[0] = end1; J.out.sec_start[0] = sec_start1; J.out.sec_width[0] = sec_width1;
        J.out.score[1] = score2; J.out.start[1] = start2; J.out.end[1] = end2; J.out.sec_start[1] = sec_start2; J.out.sec_width[1] = sec_width2;
        J.nthreads = host_threads_for(nd);
    }
    if (nd == 1) {
        jobs[0].run();
    } else {
        std::vector<std::thread> pool;
        for (int d = 0; d < nd; ++d) pool.emplace_back([&jobs, d] { jobs[d].run(); });
        for (auto& t : pool) t.join();
    }
    for (int d = 0; d < nd; ++d) {
        if (!jobs[d].cuda_error.empty()) return fail(jobs[d].cuda_error);
    }
    /* error precedence of the four reference calls, in their order (R/adaptorAlign.R:186-189) */
    FirstError ef, eb;
    for (int d = 0; d < nd; ++d) { ef.merge(jobs[d].err_front); eb.merge(jobs[d].err_back); }
    ReadView VF{front}, VB{back};
    if (tolerance > 0) {
        VF.tol = tolerance;
        VB = ReadView{front, tolerance, true};
    }
    const FirstError* base[4] = {&ef, &eb, &eb, &ef};
    const ReadView* views[4] = {&VF, &VB, &VB, &VF};
    const Plan* plans[4] = {&P[0], &P[1], &P[0], &P[1]};
    for (int r = 0; r < 4; ++r) {
        FirstError e = *base[r];
        apply_reference_errors(e, *views[r], n, *plans[r], 1);
        if (e.kind != ERR_NONE) return fail(err_text(e.kind));
    }
    return 0;
}

int sarlacc_adaptor_align_windows(const sarlacc_reads* front, const sarlacc_reads* back, const sarlacc_encoding* encoding,
        double gapopen, double gapext, const char* adaptor1, const char* adaptor2,
        int nsec1, const int32_t* sec_starts1, const int32_t* sec_ends1,
        int nsec2, const int32_t* sec_starts2, const int32_t* sec_ends2,
        const int32_t* read_width, uint8_t* reversed,
        double* score1, int32_t* start1, int32_t* end1, int32_t* sec_start1, int32_t* sec_width1,
        double* score2, int32_t* start2, int32_t* end2, int32_t* sec_start2, int32_t* sec_width2)
{
    return align_pair(front, back, 0, encoding, gapopen, gapext, adaptor1, adaptor2, nsec1, sec_starts1, sec_ends1,
                      nsec2, sec_starts2, sec_ends2, read_width, nullptr, reversed,
                      score1, start1, end1, sec_start1, sec_width1, score2, start2, end2, sec_start2, sec_width2);
}

int sarlacc_adaptor_align_reads(const sarlacc_reads* reads, int tolerance, const sarlacc_encoding* encoding,
        double gapopen, double gapext, const char* adaptor1, const char* adaptor2,
        int nsec1, const int32_t* sec_starts1, const int32_t* sec_ends1,
        int nsec2, const int32_t* sec_starts2, const int32_t* sec_ends2,
        int32_t* read_width, uint8_t* reversed,
        double* score1, int32_t* start1, int32_t* end1, int32_t* sec_start1, int32_t* sec_width1,
        double* score2, int32_t* start2, int32_t* end2, int32_t* sec_start2, int32_t* sec_width2)
{
    if (tolerance <= 0) return fail("tolerance should be a positive integer");
    return align_pair(reads, reads, tolerance, encoding, gapopen, gapext, adaptor1, adaptor2, nsec1, sec_starts1, sec_ends1,
                      nsec2, sec_starts2, sec_ends2, nullptr, read_width, reversed,
                      score1, start1, end1, sec_start1, sec_width1, score2, start2, end2, sec_start2, sec_width2);
}

int sarlacc_pack_rows(const sarlacc_reads* reads, const sarlacc_encoding* encoding, int tolerance, int back,
                      int stride, uint16_t* rows, int32_t* lens, int force_scalar)
{
    if (!reads || !rows || !lens) return fail("reads must not be NULL");
    Encoding enc;
    const char* msg = build_encoding(encoding, enc);
    if (msg) return fail(msg);
    ReadView V{reads, tolerance > 0 ? tolerance : 0, tolerance > 0 && back != 0};
    PackTables T;
    build_pack_tables(T, reads->seq_encoding, enc);
    FirstError err;
    int maxlen = 0;
    const int nt = host_threads_for(1);
    scan_lengths(V, 0, reads->n, lens, nt, err, maxlen);
    if (maxlen > stride) return fail("stride is shorter than the longest window");
    pack_rows(V, 0, reads->n, T, lens, stride, rows, nt, true, err, force_scalar != 0);
    if (err.kind != ERR_NONE) return fail(err_text(err.kind));
    return 0;
}

int sarlacc_pack_bases(const uint8_t* seq, int64_t n, int seq_encoding, uint8_t* out, int force_scalar) {
    if (n < 0 || (n > 0 && (!seq || !out))) return fail("sequence bytes must not be NULL");
    Encoding enc;
    enc.offset = 33;
    enc.n = 1;
    PackTables T;
    build_pack_tables(T, seq_encoding, enc);
    if (force_scalar) {
        for (int64_t i = 0; i + 1 < n; i += 2) out[i / 2] = (uint8_t)(T.base[seq[i]] | (T.base[seq[i + 1]] << 4));
        if (n & 1) out[n / 2] = T.base[seq[n - 1]];
    } else {
        nibble_pack(seq, (size_t)n, out, T, host_threads_for(1));
    }
    return 0;
}

/* ---- FASTQ ingest ----------------------------------------------------------------
 * Stands in for ShortRead::FastqStreamer + .FASTQ2QSDS (R/adaptorAlign.R:26,36,104-110) on the host side of the
 * ABI: a buffered reader of 4-line FASTQ records that yields chunks of reads as CSR pools (names, sequences,
 * qualities), ready to be passed back in as a sarlacc_reads.  Plain text only; host code, excluded from timings. */
struct sarlacc_fastq {
    FILE* fh = nullptr;
    gzFile gz = nullptr;          /* gzip-compressed input (ShortRead reads .gz transparently, R/adaptorAlign.R:26): inflated by zlib as it is read */
    std::vector<char> buf;
    size_t pos = 0, end = 0;
    bool eof = false;
    std::vector<uint8_t> seq_pool, qual_pool, name_pool;
    std::vector<int64_t> seq_off, qual_off, name_off;
    /* parallel condensed ingest (sarlacc_fastq_next_condensed): the file mapped read-only */
    const char* map = nullptr;
    size_t map_size = 0, map_pos = 0;
    double bytes_per_record = 0;
    std::vector<int32_t> width;

    bool fill() {
        if (eof) return false;
        if (pos > 0) {
            std::memmove(buf.data(), buf.data() + pos, end - pos);
            end -= pos;
            pos = 0;
        }
        if (end == buf.size()) buf.resize(buf.size() * 2);
        size_t got = 0;
        if (gz) {
            const int r = gzread(gz, buf.data() + end, (unsigned)std::min<size_t>(buf.size() - end, 1u << 30));
            got = r > 0 ? (size_t)r : 0;
        } else {
            got = std::fread(buf.data() + end, 1, buf.size() - end, fh);
        }
        if (got == 0) { eof = true; return false; }
        end += got;
        return true;
    }
    /* next line [b, e) without the terminator; false at end of file */
    bool line(size_t& b, size_t& e) {
        for (;;) {
            const char* nl = (const char*)std::memchr(buf.data() + pos, '\n', end - pos);
            if (nl) {
                b = pos;
                e = (size_t)(nl - buf.data());
                pos = e + 1;
                if (e > b && buf[e - 1] == '\r') --e;
                return true;
            }
            if (!fill()) {
                if (pos < end) {   /* last line without newline */
                    b = pos;
                    e = end;
                    pos = end;
                    return true;
                }
                return false;
            }
        }
    }
};

sarlacc_fastq* sarlacc_fastq_open(const char* path) {
    if (!path) { fail("path must not be NULL"); return nullptr; }
    FILE* fh = std::fopen(path, "rb");
    if (!fh) { fail(std::string("cannot open FASTQ file: ") + path); return nullptr; }
    sarlacc_fastq* f = new sarlacc_fastq();
    f->fh = fh;
    f->buf.resize(1 << 22);
    unsigned char magic[2] = {0, 0};
    const size_t got = std::fread(magic, 1, 2, fh);
    std::rewind(fh);
    if (got == 2 && magic[0] == 0x1f && magic[1] == 0x8b) {
        f->gz = gzopen(path, "rb");
        if (!f->gz) {
            std::fclose(fh);
            delete f;
            fail(std::string("cannot open gzip-compressed FASTQ file: ") + path);
            return nullptr;
        }
        gzbuffer(f->gz, 1u << 20);
    }
    return f;
}

void sarlacc_fastq_close(sarlacc_fastq* f) {
    if (!f) return;
    if (f->map && f->map_size) munmap(const_cast<char*>(f->map), f->map_size);
    if (f->gz) gzclose(f->gz);
    if (f->fh) std::fclose(f->fh);
    delete f;
}

/* Reads up to max_reads records.  Returns the number read (0 at end of file), or -1 on a malformed record.  The
 * pointers stay valid until the next call on this handle. */
int64_t sarlacc_fastq_next(sarlacc_fastq* f, int64_t max_reads,
        const uint8_t** seq_pool, const int64_t** seq_off, const uint8_t** qual_pool, const int64_t** qual_off,
        const uint8_t** name_pool, const int64_t** name_off)
{
    if (!f) { fail("FASTQ handle is NULL"); return -1; }
    f->seq_pool.clear(); f->qual_pool.clear(); f->name_pool.clear();
    f->seq_off.assign(1, 0); f->qual_off.assign(1, 0); f->name_off.assign(1, 0);
    int64_t n = 0;
    size_t b, e;
    while (n < max_reads) {
        if (!f->line(b, e)) break;
        if (e == b) continue;                       /* blank line between records */
        if (f->buf[b] != '@') { fail("malformed FASTQ record: header does not start with '@'"); return -1; }
        f->name_pool.insert(f->name_pool.end(), f->buf.begin() + b + 1, f->buf.begin() + e);
        f->name_off.push_back((int64_t)f->name_pool.size());
        if (!f->line(b, e)) { fail("malformed FASTQ record: missing sequence line"); return -1; }
        f->seq_pool.insert(f->seq_pool.end(), f->buf.begin() + b, f->buf.begin() + e);
        f->seq_off.push_back((int64_t)f->seq_pool.size());
        if (!f->line(b, e) || e == b || f->buf[b] != '+') { fail("malformed FASTQ record: missing '+' line"); return -1; }
        if (!f->line(b, e)) { fail("malformed FASTQ record: missing quality line"); return -1; }
        f->qual_pool.insert(f->qual_pool.end(), f->buf.begin() + b, f->buf.begin() + e);
        f->qual_off.push_back((int64_t)f->qual_pool.size());
        ++n;
    }
    if (f->seq_pool.empty()) f->seq_pool.push_back(0);
    if (f->qual_pool.empty()) f->qual_pool.push_back(0);
    if (f->name_pool.empty()) f->name_pool.push_back(0);
    if (seq_pool) *seq_pool = f->seq_pool.data();
    if (seq_off) *seq_off = f->seq_off.data();
    if (qual_pool) *qual_pool = f->qual_pool.data();
    if (qual_off) *qual_off = f->qual_off.data();
    if (name_pool) *name_pool = f->name_pool.data();
    if (name_off) *name_off = f->name_off.data();
    return n;
}

/* ---- parallel condensed ingest ------------------------------------------------------------------
 * adaptorAlign only ever looks at the first and last `tolerance` bases of a read (R/adaptorAlign.R:86-95), so the
 * ingest that feeds it keeps just those (plus name and length): ~0.5 KB instead of ~10 KB per 5 kb read.  The mapped
 * file is cut into one byte range per thread; a range starts at the first line that begins with '@' AND whose second
 * successor begins with '+' -- a quality line may start with '@', but then the line two below it is a sequence line,
 * which cannot start with '+'. */
namespace {

struct FqPiece {
    std::vector<uint8_t> seq, qual, name;
    std::vector<int32_t> seq_len, name_len, width;
    std::vector<size_t> rec_end;     /* byte after each record */
    std::string error;
    bool complete = false;           /* the whole byte range was parsed (not stopped by the record cap) */
};

inline const char* fq_line_end(const char* p, const char* end) {
    const char* nl = (const char*)std::memchr(p, '\n', (size_t)(end - p));
    return nl ? nl : end;
}

/* first record start at or after `from` (which must be a line start or get aligned to one) */
const char* fq_record_start(const char* base, const char* from, const char* end, bool align) {
    const char* p = from;
    if (align && p > base && p[-1] != '\n') {
        p = fq_line_end(p, end);
        if (p < end) ++p;
    }
    while (p < end) {
        const char* e1 = fq_line_end(p, end);
        if (*p == '@') {
            const char* l2 = e1 < end ? e1 + 1 : end;
            const char* e2 = fq_line_end(l2, end);
            const char* l3 = e2 < end ? e2 + 1 : end;
            if (l3 < end && *l3 == '+') return p;
        }
        p = e1 < end ? e1 + 1 : end;
    }
    return end;
}

void fq_parse_range(const char* base, const char* p, const char* stop, const char* end, int keep, int64_t max_records, FqPiece& out) {
    auto trimmed = [](const char* b, const char* e) { return (e > b && e[-1] == '\r') ? e - 1 : e; };
    int64_t n = 0;
    while (p < stop && n < max_records) {
        const char* e = fq_line_end(p, end);
        if (trimmed(p, e) == p) { p = e < end ? e + 1 : end; continue; }          /* blank line between records */
        if (*p != '@') { out.error = "malformed FASTQ record: header does not start with '@'"; return; }
        const char* hb = p + 1;
        const char* he = trimmed(p, e);
        if (e >= end) { out.error = "malformed FASTQ record: missing sequence line"; return; }
        const char* sb = e + 1;
        const char* se_raw = fq_line_end(sb, end);
        const char* se = trimmed(sb, se_raw);
        if (se_raw >= end) { out.error = "malformed FASTQ record: missing '+' line"; return; }
        const char* pb = se_raw + 1;
        const char* pe = fq_line_end(pb, end);
        if (pb >= end || *pb != '+') { out.error = "malformed FASTQ record: missing '+' line"; return; }
        if (pe >= end) { out.error = "malformed FASTQ record: missing quality line"; return; }
        const char* qb = pe + 1;
        const char* qe_raw = fq_line_end(qb, end);
        const char* qe = trimmed(qb, qe_raw);
        const int64_t L = se - sb;
        if (qe - qb != L) { out.error = "malformed FASTQ record: sequence and quality lengths differ"; return; }
        if (L > 0x7fffffff) { out.error = "malformed FASTQ record: read longer than 2^31 bases"; return; }
        out.name.insert(out.name.end(), hb, he);
        out.name_len.push_back((int32_t)(he - hb));
        if (L <= 2 * (int64_t)keep) {
            out.seq.insert(out.seq.end(), sb, se);
            out.qual.insert(out.qual.end(), qb, qe);
            out.seq_len.push_back((int32_t)L);
        } else {
            out.seq.insert(out.seq.end(), sb, sb + keep);
            out.seq.insert(out.seq.end(), se - keep, se);
            out.qual.insert(out.qual.end(), qb, qb + keep);
            out.qual.insert(out.qual.end(), qe - keep, qe);
            out.seq_len.push_back(2 * keep);
        }
        out.width.push_back((int32_t)L);
        p = qe_raw < end ? qe_raw + 1 : end;
        out.rec_end.push_back((size_t)(p - base));
        ++n;
    }
    out.complete = p >= stop;
}

}  // namespace

int64_t sarlacc_fastq_next_condensed(sarlacc_fastq* f, int64_t max_reads, int keep, int nthreads,
        const uint8_t** seq_pool, const int64_t** seq_off, const uint8_t** qual_pool, const int64_t** qual_off,
        const uint8_t** name_pool, const int64_t** name_off, const int32_t** width)
{
    if (!f) { fail("FASTQ handle is NULL"); return -1; }
    if (keep < 1) { fail("the number of bases to keep per read end must be positive"); return -1; }
    if (max_reads < 1) max_reads = 1;
    if (f->gz) {
        /* a gzip stream cannot be cut into ranges: records are parsed one after the other as zlib inflates them, and
         * condensed on the fly (first and last `keep` bases; the whole read when it is no longer than 2 * keep) */
        f->seq_pool.clear(); f->qual_pool.clear(); f->name_pool.clear(); f->width.clear();
        f->seq_off.assign(1, 0); f->qual_off.assign(1, 0); f->name_off.assign(1, 0);
        int64_t n = 0;
        size_t b, e;
        auto condense = [&](std::vector<uint8_t>& pool, size_t lo, size_t hi) {
            const size_t len = hi - lo;
            if (len <= 2 * (size_t)keep) {
                pool.insert(pool.end(), f->buf.begin() + lo, f->buf.begin() + hi);
            } else {
                pool.insert(pool.end(), f->buf.begin() + lo, f->buf.begin() + lo + keep);
                pool.insert(pool.end(), f->buf.begin() + hi - keep, f->buf.begin() + hi);
            }
        };
        while (n < max_reads) {
            if (!f->line(b, e)) break;
            if (e == b) continue;
            if (f->buf[b] != '@') { fail("malformed FASTQ record: header does not start with '@'"); return -1; }
            f->name_pool.insert(f->name_pool.end(), f->buf.begin() + b + 1, f->buf.begin() + e);
            f->name_off.push_back((int64_t)f->name_pool.size());
            if (!f->line(b, e)) { fail("malformed FASTQ record: missing sequence line"); return -1; }
            const size_t slen = e - b;
            condense(f->seq_pool, b, e);
            f->seq_off.push_back((int64_t)f->seq_pool.size());
            if (!f->line(b, e) || e == b || f->buf[b] != '+') { fail("malformed FASTQ record: missing '+' line"); return -1; }
            if (!f->line(b, e)) { fail("malformed FASTQ record: missing quality line"); return -1; }
            if (e - b != slen) { fail("malformed FASTQ record: sequence and quality lengths differ"); return -1; }
            condense(f->qual_pool, b, e);
            f->qual_off.push_back((int64_t)f->qual_pool.size());
            f->width.push_back((int32_t)slen);
            ++n;
        }
        if (f->seq_pool.empty()) f->seq_pool.push_back(0);
        if (f->qual_pool.empty()) f->qual_pool.push_back(0);
        if (f->name_pool.empty()) f->name_pool.push_back(0);
        if (f->width.empty()) f->width.push_back(0);
        if (seq_pool) *seq_pool = f->seq_pool.data();
        if (seq_off) *seq_off = f->seq_off.data();
        if (qual_pool) *qual_pool = f->qual_pool.data();
        if (qual_off) *qual_off = f->qual_off.data();
        if (name_pool) *name_pool = f->name_pool.data();
        if (name_off) *name_off = f->name_off.data();
        if (width) *width = f->width.data();
        return n;
    }
    if (!f->map) {
        if (fseeko(f->fh, 0, SEEK_END) != 0) { fail("cannot seek in FASTQ file"); return -1; }
        const off_t sz = ftello(f->fh);
        f->map_size = (size_t)sz;
        if (sz > 0) {
            void* m = mmap(nullptr, (size_t)sz, PROT_READ, MAP_PRIVATE, fileno(f->fh), 0);
            if (m == MAP_FAILED) { f->map_size = 0; fail("cannot map FASTQ file"); return -1; }
            f->map = (const char*)m;
            madvise(m, (size_t)sz, MADV_SEQUENTIAL);
        } else {
            f->map = "";
        }
    }
    if (nthreads <= 0) nthreads = host_threads_for(1);
    f->seq_pool.clear(); f->qual_pool.clear(); f->name_pool.clear(); f->width.clear();
    f->seq_off.assign(1, 0); f->qual_off.assign(1, 0); f->name_off.assign(1, 0);
    const char* base = f->map;
    const char* end = base + f->map_size;
    int64_t got = 0;
    while (got < max_reads && f->map_pos < f->map_size) {
        const int64_t want = max_reads - got;
        /* byte budget for this round: the records still wanted at the size seen so far (+5 %), at least 1 MiB per thread */
        double bpr = f->bytes_per_record > 0 ? f->bytes_per_record : 4096.0;
        size_t budget = (size_t)std::min<double>((double)(f->map_size - f->map_pos), std::max(1.05 * bpr * (double)want, (double)nthreads * (1 << 20)));
        const char* lo = base + f->map_pos;
        const char* hi = lo + budget;
        int T = (int)std::min<size_t>((size_t)nthreads, std::max<size_t>(1, budget >> 20));
        std::vector<const char*> cut((size_t)T + 1);
        cut[0] = lo;
        for (int t = 1; t < T; ++t) cut[(size_t)t] = fq_record_start(base, lo + budget / T * t, end, true);
        cut[(size_t)T] = (hi >= end) ? end : fq_record_start(base, hi, end, true);
        for (int t = 1; t <= T; ++t) cut[(size_t)t] = std::max(cut[(size_t)t], cut[(size_t)t - 1]);
        std::vector<FqPiece> pieces((size_t)T);
        /* (reading the ranges with pread into private buffers instead of parsing the mapping was 3x slower here) */
        const std::function<void(int)> task = [&](int t) {
            fq_parse_range(base, cut[(size_t)t], cut[(size_t)t + 1], end, keep, t == 0 ? want : (int64_t)1 << 62, pieces[(size_t)t]);
        };
        if (T == 1) task(0); else WorkerPool::instance().run(T, task);
        size_t consumed_to = f->map_pos;
        bool stop = false;
        for (int t = 0; t < T && !stop; ++t) {
            FqPiece& P = pieces[(size_t)t];
            size_t so = 0, no = 0;
            for (size_t r = 0; r < P.width.size(); ++r) {
                if (got >= max_reads) { stop = true; break; }
                const size_t sl = (size_t)P.seq_len[r], nl = (size_t)P.name_len[r];
                f->seq_pool.insert(f->seq_pool.end(), P.seq.begin() + so, P.seq.begin() + so + sl);
                f->qual_pool.insert(f->qual_pool.end(), P.qual.begin() + so, P.qual.begin() + so + sl);
                f->name_pool.insert(f->name_pool.end(), P.name.begin() + no, P.name.begin() + no + nl);
                so += sl;
                no += nl;
                f->seq_off.push_back((int64_t)f->seq_pool.size());
                f->qual_off.push_back((int64_t)f->qual_pool.size());
                f->name_off.push_back((int64_t)f->name_pool.size());
                f->width.push_back(P.width[r]);
                consumed_to = P.rec_end[r];
                ++got;
            }
            if (stop) break;
            if (!P.error.empty()) { fail(P.error); return -1; }
            if (P.complete) consumed_to = std::max(consumed_to, (size_t)(cut[(size_t)t + 1] - base));   /* incl. trailing blank lines */
            else stop = true;      /* stopped at its record cap: the next piece does not follow on */
        }
        const size_t before = f->map_pos;
        f->map_pos = consumed_to;
        if (got > 0 && f->map_pos > before) f->bytes_per_record = (double)(f->map_pos - before) / (double)std::max<int64_t>(1, got);
        if (f->map_pos == before) break;     /* no progress: trailing bytes without a record */
    }
    if (f->seq_pool.empty()) f->seq_pool.push_back(0);
    if (f->qual_pool.empty()) f->qual_pool.push_back(0);
    if (f->name_pool.empty()) f->name_pool.push_back(0);
    if (f->width.empty()) f->width.push_back(0);
    if (seq_pool) *seq_pool = f->seq_pool.data();
    if (seq_off) *seq_off = f->seq_off.data();
    if (qual_pool) *qual_pool = f->qual_pool.data();
    if (qual_off) *qual_off = f->qual_off.data();
    if (name_pool) *name_pool = f->name_pool.data();
    if (name_off) *name_off = f->name_off.data();
    if (width) *width = f->width.data();
    return got;
}

/* ---- resident windows ------------------------------------------------------------------------- */

struct sarlacc_resident {
    int device = 0;
    int64_t n = 0;
    int stride = 0;
    int maxlen = 0;
    int64_t total_len = 0;
    Encoding enc;
    cudaStream_t own_stream = nullptr;
    cudaStream_t tb_stream = nullptr;            /* tracebacks of sub-range k overlap the forward pass of k+1 */
    cudaEvent_t fwd_done[2] = {nullptr, nullptr};
    cudaEvent_t tb_done[2] = {nullptr, nullptr};
    bool tb_pending[2] = {false, false};
    DevBuf d_rows, d_lens, d_out;
    Scratch scratch, scratch2;
    /* plans (host tables + their device copy) are cached per (reference, penalties, mode, sections) so that the
     * steady state of adaptorAlign -> getAdaptorThresholds style re-use enqueues kernels only */
    struct CachedPlan {
        std::string key;
        Plan plan;
        DevPlan d;
    };
    std::vector<std::unique_ptr<CachedPlan> > plans;
    CachedPlan* cur = nullptr;
    OutLayout lay;
    int nsec = 0;
    Mode mode = MODE_SCORE_LOCAL;
    bool has_result = false;
    int sms = 0;
    std::string last_kernel;
    std::vector<int32_t> h_lens;
    FwdTimer timer;
    bool timing = false;
    long long sub = 0;     /* alignments per sub-range of the last run */
};

sarlacc_resident* sarlacc_resident_create(const sarlacc_reads* reads, const sarlacc_encoding* encoding, int device) {
    if (!reads) { fail("reads must not be NULL"); return nullptr; }
    if (require_device()) return nullptr;
    std::unique_ptr<sarlacc_resident> r(new sarlacc_resident());
    const char* msg = build_encoding(encoding, r->enc);
    if (msg) { fail(msg); return nullptr; }
    try {
        CUDA_CHECK(cudaSetDevice(device));
        r->device = device;
        r->sms = device_sm_count(device);
        r->n = reads->n;
        CUDA_CHECK(cudaStreamCreateWithFlags(&r->own_stream, cudaStreamNonBlocking));
        CUDA_CHECK(cudaStreamCreateWithFlags(&r->tb_stream, cudaStreamNonBlocking));
        for (int k = 0; k < 2; ++k) {
            CUDA_CHECK(cudaEventCreateWithFlags(&r->fwd_done[k], cudaEventDisableTiming));
            CUDA_CHECK(cudaEventCreateWithFlags(&r->tb_done[k], cudaEventDisableTiming));
        }
        ReadView V{reads};
        PackTables PT;
        build_pack_tables(PT, reads->seq_encoding, r->enc);
        const int nthreads = host_threads_for(1);
        r->h_lens.assign((size_t)std::max<int64_t>(r->n, 1), 0);
        FirstError err;
        scan_lengths(V, 0, r->n, r->h_lens.data(), nthreads, err, r->maxlen);
        r->stride = std::max(8, (r->maxlen + 8) & ~7);
        r->total_len = 0;
        for (int64_t i = 0; i < r->n; ++i) r->total_len += r->h_lens[(size_t)i];
        /* pack + upload in pieces through a pinned staging buffer */
        const int64_t piece = std::max<int64_t>(1, (int64_t)(64u << 20) / ((int64_t)r->stride * 2));
        PinBuf stage[2];
        cudaEvent_t ev[2];
        CUDA_CHECK(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
        bool used[2] = {false, false};
        r->d_rows.reserve(sizeof(uint16_t) * (size_t)std::max<int64_t>(r->n, 1) * r->stride);
        r->d_lens.reserve(sizeof(int32_t) * (size_t)std::max<int64_t>(r->n, 1));
        int w = 0;
        for (int64_t c0 = 0; c0 < r->n; c0 += piece, w ^= 1) {
            const int64_t c1 = std::min(r->n, c0 + piece);
            if (used[w]) CUDA_CHECK(cudaEventSynchronize(ev[w]));
            stage[w].reserve(sizeof(uint16_t) * (size_t)(c1 - c0) * r->stride);
            pack_rows(V, c0, c1, PT, r->h_lens.data() + c0, r->stride, stage[w].as<uint16_t>(), nthreads, true, err);
            CUDA_CHECK(cudaMemcpyAsync(r->d_rows.as<uint16_t>() + (size_t)c0 * r->stride, stage[w].p,
                                       sizeof(uint16_t) * (size_t)(c1 - c0) * r->stride, cudaMemcpyHostToDevice, r->own_stream));
            CUDA_CHECK(cudaEventRecord(ev[w], r->own_stream));
            used[w] = true;
        }
        if (r->n > 0) {
            CUDA_CHECK(cudaMemcpyAsync(r->d_lens.p, r->h_lens.data(), sizeof(int32_t) * (size_t)r->n, cudaMemcpyHostToDevice, r->own_stream));
        }
        CUDA_CHECK(cudaStreamSynchronize(r->own_stream));
        cudaEventDestroy(ev[0]);
        cudaEventDestroy(ev[1]);
        stage[0].release();
        stage[1].release();
        if (err.kind != ERR_NONE) {
            fail(err_text(err.kind));
            sarlacc_resident_free(r.release());
            return nullptr;
        }
    } catch (CudaError& e) {
        fail(e.msg);
        sarlacc_resident_free(r.release());
        return nullptr;
    }
    return r.release();
}

sarlacc_resident* sarlacc_resident_scrambled(const sarlacc_resident* src, uint64_t seed, uint64_t first_index,
        const uint64_t* read_index, int stream_id)
{
    if (!src) { fail("resident handle is NULL"); return nullptr; }
    std::unique_ptr<sarlacc_resident> r(new sarlacc_resident());
    try {
        CUDA_CHECK(cudaSetDevice(src->device));
        r->device = src->device;
        r->sms = src->sms;
        r->n = src->n;
        r->stride = src->stride;
        r->maxlen = src->maxlen;
        r->total_len = src->total_len;
        r->enc = src->enc;
        r->h_lens = src->h_lens;
        CUDA_CHECK(cudaStreamCreateWithFlags(&r->own_stream, cudaStreamNonBlocking));
        CUDA_CHECK(cudaStreamCreateWithFlags(&r->tb_stream, cudaStreamNonBlocking));
        for (int k = 0; k < 2; ++k) {
            CUDA_CHECK(cudaEventCreateWithFlags(&r->fwd_done[k], cudaEventDisableTiming));
            CUDA_CHECK(cudaEventCreateWithFlags(&r->tb_done[k], cudaEventDisableTiming));
        }
        const size_t nn = (size_t)std::max<int64_t>(r->n, 1);
        r->d_rows.reserve(sizeof(uint16_t) * nn * r->stride);
        r->d_lens.reserve(sizeof(int32_t) * nn);
        CUDA_CHECK(cudaMemsetAsync(r->d_rows.p, 0, sizeof(uint16_t) * nn * r->stride, r->own_stream));
        if (r->n > 0) {
            CUDA_CHECK(cudaMemcpyAsync(r->d_lens.p, src->d_lens.p, sizeof(int32_t) * (size_t)r->n, cudaMemcpyDeviceToDevice, r->own_stream));
        }
        DevBuf idx;
        if (read_index && r->n > 0) {
            idx.reserve(sizeof(uint64_t) * (size_t)r->n);
            CUDA_CHECK(cudaMemcpyAsync(idx.p, read_index, sizeof(uint64_t) * (size_t)r->n, cudaMemcpyHostToDevice, r->own_stream));
        }
        /* the source may still be busy on its own stream */
        CUDA_CHECK(cudaStreamSynchronize(src->own_stream));
        launch_scramble(src->d_rows.as<uint16_t>(), r->d_rows.as<uint16_t>(), r->d_lens.as<int32_t>(), r->n, r->stride,
                        seed, first_index, read_index ? idx.as<unsigned long long>() : nullptr, (unsigned long long)stream_id, r->own_stream);
        g_launches += 1;
        CUDA_CHECK(cudaGetLastError());
        CUDA_CHECK(cudaStreamSynchronize(r->own_stream));
        idx.release();
    } catch (CudaError& e) {
        fail(e.msg);
        sarlacc_resident_free(r.release());
        return nullptr;
    }
    return r.release();
}

int sarlacc_resident_rows(sarlacc_resident* r, uint16_t* rows, int32_t* lens, int* stride) {
    if (!r) return fail("resident handle is NULL");
    try {
        CUDA_CHECK(cudaSetDevice(r->device));
        if (stride) *stride = r->stride;
        if (r->n > 0 && rows) CUDA_CHECK(cudaMemcpy(rows, r->d_rows.p, sizeof(uint16_t) * (size_t)r->n * r->stride, cudaMemcpyDeviceToHost));
        if (r->n > 0 && lens) CUDA_CHECK(cudaMemcpy(lens, r->d_lens.p, sizeof(int32_t) * (size_t)r->n, cudaMemcpyDeviceToHost));
    } catch (CudaError& e) {
        return fail(e.msg);
    }
    return 0;
}

void sarlacc_resident_free(sarlacc_resident* r) {
    if (!r) return;
    cudaSetDevice(r->device);
    r->d_rows.release();
    r->d_lens.release();
    r->d_out.release();
    r->scratch.release();
    r->scratch2.release();
    for (int k = 0; k < 2; ++k) {
        if (r->fwd_done[k]) cudaEventDestroy(r->fwd_done[k]);
        if (r->tb_done[k]) cudaEventDestroy(r->tb_done[k]);
    }
    if (r->tb_stream) cudaStreamDestroy(r->tb_stream);
    for (auto& p : r->plans) p->d.buf.release();
    r->plans.clear();
    r->timer.release();
    if (r->own_stream) cudaStreamDestroy(r->own_stream);
    delete r;
}

int64_t sarlacc_resident_n(const sarlacc_resident* r) { return r ? r->n : 0; }

int64_t sarlacc_resident_cells(const sarlacc_resident* r, int rlen) { return r ? r->total_len * (int64_t)rlen : 0; }

int64_t sarlacc_resident_bytes(const sarlacc_resident* r) {
    return r ? (int64_t)sizeof(uint16_t) * r->n * r->stride + (int64_t)sizeof(int32_t) * r->n : 0;
}

int sarlacc_resident_align(sarlacc_resident* r, int mode, double gapopen, double gapext, const char* reference,
        int nsec, const int32_t* sec_starts, const int32_t* sec_ends, void* stream)
{
    if (!r) return fail("resident handle is NULL");
    if (!reference) return fail("adaptor sequence should be a string");
    if (mode < 0 || mode > 2) return fail("mode must be 0 (score, local), 1 (traceback, local) or 2 (score, global)");
    const Mode md = (Mode)mode;
    const bool local = md != MODE_SCORE_GLOBAL;
    const bool trace = md == MODE_TRACE_LOCAL;
    const int L = (int)std::strlen(reference);
    if (L == 0) return fail("resident runs need a non-empty reference");
    try {
        CUDA_CHECK(cudaSetDevice(r->device));
        cudaStream_t st = stream ? reinterpret_cast<cudaStream_t>(stream) : r->own_stream;
        std::string key(reference);
        key += '|';
        key.append(reinterpret_cast<const char*>(&gapopen), sizeof(double));
        key.append(reinterpret_cast<const char*>(&gapext), sizeof(double));
        key += local ? 'L' : 'G';
        key += trace ? 'T' : 'S';
        if (trace) {
            if (nsec < 0 || (nsec > 0 && (!sec_starts || !sec_ends))) return fail("section starts and ends should have the same length");
            key.append(reinterpret_cast<const char*>(sec_starts), sizeof(int32_t) * nsec);
            key += '|';
            key.append(reinterpret_cast<const char*>(sec_ends), sizeof(int32_t) * nsec);
        }
        sarlacc_resident::CachedPlan* cp = nullptr;
        for (auto& p : r->plans) {
            if (p->key == key) { cp = p.get(); break; }
        }
        if (!cp) {
            if (r->plans.size() >= 64) {   /* bounded: drop everything once the device is idle */
                CUDA_CHECK(cudaStreamSynchronize(st));
                for (auto& p : r->plans) p->d.buf.release();
                r->plans.clear();
            }
            std::unique_ptr<sarlacc_resident::CachedPlan> np(new sarlacc_resident::CachedPlan());
            np->key = key;
            const char* refs[1] = {reference};
            build_plan(np->plan, r->enc, refs, 1, L, local, gapopen, gapext, trace);
            if (trace && nsec > 0) {
                np->plan.sec_starts.assign(sec_starts, sec_starts + nsec);
                np->plan.sec_ends.assign(sec_ends, sec_ends + nsec);
                for (int s = 0; s < nsec; ++s) {
                    if (sec_starts[s] < 0 || sec_starts[s] > L || sec_ends[s] < 0 || sec_ends[s] > L) return fail("section bounds outside the adaptor");
                }
            }
            np->d.upload(np->plan, st);
            cp = np.get();
            r->plans.push_back(std::move(np));
        }
        if (cp->plan.bad_col[0] >= 0 && r->total_len > 0) return fail(err_text(ERR_REF));
        r->cur = cp;
        r->nsec = trace ? nsec : 0;
        r->mode = md;
        r->lay = make_layout(md, std::max<int64_t>(r->n, 1), 1, r->nsec, r->maxlen, L);
        r->d_out.reserve(r->lay.total);
        uint8_t* d = r->d_out.as<uint8_t>();
        /* Sub-chunks (scratch budget) write straight into the full-size output block: sections are laid out
         * [nsec][n] so each sub-chunk gets its own launch set with per-section base pointers. */
        /* two scratch sets: each sub-range gets half of the budget so that traceback(k) and forward(k+1) overlap */
        long long cn = sub_chunk(cp->plan, r->maxlen, trace, std::max<int64_t>(r->n, 1));
        const bool overlap = trace && std::getenv("SARLACC_NO_OVERLAP") == nullptr;
        if (overlap && cn >= r->n && r->n >= 65536) cn = (r->n + 1) / 2;          /* at least two ranges to overlap */
        else if (overlap && cn < r->n) cn = std::max<long long>(1, cn / 2);
        if (cn < r->n) cn = whole_rounds(cn, plan_groups(cp->plan, trace));
        r->sub = cn;
        const char* name = "";
        r->timer.used = 0;
        int k = 0;
        for (long long off = 0; off < r->n; off += cn, ++k) {
            const long long m = std::min<long long>(cn, r->n - off);
            const int b = k & 1;
            Outputs dev;
            dev.score = reinterpret_cast<double*>(d + r->lay.o_score) + off;
            if (trace) {
                dev.start = reinterpret_cast<int32_t*>(d + r->lay.o_start) + off;
                dev.end = reinterpret_cast<int32_t*>(d + r->lay.o_end) + off;
                /* section matrices: one run -> [nsec][n]; split run -> consecutive compact [nsec][m] blocks,
                 * scattered by sarlacc_resident_fetch */
                dev.sec_start = reinterpret_cast<int32_t*>(d + r->lay.o_ss) + (cn >= r->n ? 0 : off * (long long)std::max(1, r->nsec));
                dev.sec_width = reinterpret_cast<int32_t*>(d + r->lay.o_sw) + (cn >= r->n ? 0 : off * (long long)std::max(1, r->nsec));
            }
            if (overlap && r->tb_pending[b]) {
                /* the forward pass of this range overwrites the records traceback(k-2) may still be reading */
                CUDA_CHECK(cudaStreamWaitEvent(st, r->tb_done[b], 0));
                r->tb_pending[b] = false;
            }
            name = run_device(cp->plan, cp->d, b ? r->scratch2 : r->scratch, st, r->d_rows.as<uint16_t>() + (size_t)off * r->stride,
                              r->d_lens.as<int32_t>() + off, m, r->stride, r->maxlen, trace, dev, r->sms,
                              r->timing ? &r->timer : nullptr,
                              overlap ? r->tb_stream : nullptr, r->fwd_done[b], r->tb_done[b]);
            if (overlap) r->tb_pending[b] = true;
        }
        /* the caller's stream sees the run as complete only when the tracebacks are */
        for (int b = 0; b < 2; ++b) {
            if (r->tb_pending[b]) {
                CUDA_CHECK(cudaStreamWaitEvent(st, r->tb_done[b], 0));
                r->tb_pending[b] = false;
            }
        }
        {
            const Geometry g = geometry_for(cp->plan, r->maxlen);
            r->last_kernel = std::string(name) + " G=" + std::to_string(g.G) + " C=" + std::to_string(g.C);
        }
        r->has_result = true;
    } catch (CudaError& e) {
        return fail(e.msg);
    }
    return 0;
}

int sarlacc_resident_fetch(sarlacc_resident* r, double* score, int32_t* start, int32_t* end,
        int32_t* sec_start, int32_t* sec_width, void* stream)
{
    if (!r) return fail("resident handle is NULL");
    if (!r->has_result) return fail("no resident run to fetch");
    try {
        CUDA_CHECK(cudaSetDevice(r->device));
        cudaStream_t st = stream ? reinterpret_cast<cudaStream_t>(stream) : r->own_stream;
        const uint8_t* d = r->d_out.as<uint8_t>();
        const size_t n = (size_t)r->n;
        if (n == 0) return 0;
        if (score) CUDA_CHECK(cudaMemcpyAsync(score, d + r->lay.o_score, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
        if (r->mode == MODE_TRACE_LOCAL) {
            if (start) CUDA_CHECK(cudaMemcpyAsync(start, d + r->lay.o_start, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
            if (end) CUDA_CHECK(cudaMemcpyAsync(end, d + r->lay.o_end, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
            const long long cn = r->sub > 0 ? r->sub : (long long)n;
            if (r->nsec > 0 && (sec_start || sec_width)) {
                if (cn >= (long long)n) {
                    if (sec_start) CUDA_CHECK(cudaMemcpyAsync(sec_start, d + r->lay.o_ss, sizeof(int32_t) * n * r->nsec, cudaMemcpyDeviceToHost, st));
                    if (sec_width) CUDA_CHECK(cudaMemcpyAsync(sec_width, d + r->lay.o_sw, sizeof(int32_t) * n * r->nsec, cudaMemcpyDeviceToHost, st));
                } else {
                    /* split run: device holds consecutive [nsec][m] blocks */
                    for (long long off = 0; off < (long long)n; off += cn) {
                        const long long m = std::min<long long>(cn, (long long)n - off);
                        for (int s = 0; s < r->nsec; ++s) {
                            const size_t src = sizeof(int32_t) * ((size_t)off * r->nsec + (size_t)s * m);
                            if (sec_start) CUDA_CHECK(cudaMemcpyAsync(sec_start + (size_t)s * n + off, d + r->lay.o_ss + src, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, st));
                            if (sec_width) CUDA_CHECK(cudaMemcpyAsync(sec_width + (size_t)s * n + off, d + r->lay.o_sw + src, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, st));
                        }
                    }
                }
            }
        }
        CUDA_CHECK(cudaStreamSynchronize(st));
    } catch (CudaError& e) {
        return fail(e.msg);
    }
    return 0;
}

const double* sarlacc_resident_scores_device(const sarlacc_resident* r) {
    if (!r || !r->has_result) return nullptr;
    return reinterpret_cast<const double*>(r->d_out.as<uint8_t>() + r->lay.o_score);
}

const char* sarlacc_resident_last_kernel(const sarlacc_resident* r) { return r ? r->last_kernel.c_str() : ""; }

void sarlacc_resident_set_timing(sarlacc_resident* r, int on) {
    if (r) r->timing = on != 0;
}

double sarlacc_resident_forward_ms(sarlacc_resident* r) {
    if (!r) return -1.0;
    try {
        CUDA_CHECK(cudaSetDevice(r->device));
        return r->timer.total_ms();
    } catch (CudaError& e) {
        fail(e.msg);
        return -1.0;
    }
}


/* ---- chunks: device-resident reads that are re-loaded in place ---------------------------------------------------
 * One chunk = what one FastqStreamer yield() is to the R drivers (R/adaptorAlign.R:26-48, R/getAdaptorThresholds.R:35-48):
 * up to `capacity` reads whose front and back windows sit packed in HBM.  Loading replaces the contents in place (no
 * allocation in steady state); sarlacc_chunk_adaptor_align is .align_AA_internal + adaptorAlign's adaptor2 flip,
 * sarlacc_chunk_scrambled_scores is .align_AT_internal.  Every call only enqueues work (compute, traceback and copy
 * streams of the chunk) and results land in caller memory -- host or device -- by the time sarlacc_chunk_sync returns,
 * so the copy-out of chunk k overlaps the alignment of chunk k+1. */
struct sarlacc_chunk {
    int device = 0, sms = 0;
    int64_t capacity = 0, n = 0, scrambled_n = -1;
    int tol = 0, stride = 0, maxlen = 0;
    bool has_width = false;
    Encoding enc;
    int seq_encoding = SARLACC_SEQ_ASCII;
    cudaStream_t st = nullptr, tb = nullptr, cp = nullptr;
    cudaStream_t fw[2] = {nullptr, nullptr};    /* forward passes of the sub-ranges of either scratch parity (see sarlacc_chunk_adaptor_align) */
    cudaEvent_t rows_ready = nullptr;
    DevBuf rows_f, rows_b, lens_f, lens_b, width, flipped, srows_f, srows_b, barcodes;
    RawStage raw_f, raw_b;
    PinBuf h_lens_f, h_lens_b, h_width;
    struct CachedPlan { std::string key; Plan plan; DevPlan d; };
    std::vector<std::unique_ptr<CachedPlan> > plans;
    PairScratch pair[2];
    Scratch score_scratch, score_scratch2;   /* one per stream of the score-only passes */
    SideStream side;
    cudaEvent_t fwd_done[2] = {nullptr, nullptr}, tb_done[2] = {nullptr, nullptr};
    bool tb_pending[2] = {false, false};
    int parity = 0;
    DevBuf tmp;                     /* [4][capacity] forward scores */
    DevBuf out[2];                  /* final columns of the last two adaptor_align calls */
    DevBuf sout[2];                 /* kept scrambled scores of the last two scrambled_scores calls */
    cudaEvent_t out_ready = nullptr, out_copied[2] = {nullptr, nullptr}, sout_copied[2] = {nullptr, nullptr};
    bool out_pending[2] = {false, false}, sout_pending[2] = {false, false};
    int out_which = 0, sout_which = 0;
    /* phase timing (sarlacc_chunk_set_timing): CUDA events on the compute stream around load / align / scramble / score */
    bool timing = false;
    std::vector<cudaEvent_t> tev;   /* pairs, tagged by phase */
    std::vector<int> tphase;
    double phase_ms[4] = {0, 0, 0, 0};
    std::string last_kernel[2];
};

namespace {

void chunk_mark(sarlacc_chunk* c, int phase, bool begin) {
    if (!c->timing) return;
    cudaEvent_t e;
    CUDA_CHECK(cudaEventCreate(&e));
    CUDA_CHECK(cudaEventRecord(e, c->st));
    c->tev.push_back(e);
    if (begin) c->tphase.push_back(phase);
}

void chunk_collect_timing(sarlacc_chunk* c) {
    for (size_t i = 0; i + 1 < c->tev.size(); i += 2) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c->tev[i], c->tev[i + 1]) == cudaSuccess) c->phase_ms[c->tphase[i / 2]] += ms;
    }
    for (auto e : c->tev) cudaEventDestroy(e);
    c->tev.clear();
    c->tphase.clear();
}

sarlacc_chunk::CachedPlan* chunk_plan(sarlacc_chunk* c, const char* reference, double go, double ge, bool trace,
                                      int nsec, const int32_t* ss, const int32_t* se) {
    std::string key(reference);
    key += '|';
    key.append(reinterpret_cast<const char*>(&go), sizeof(double));
    key.append(reinterpret_cast<const char*>(&ge), sizeof(double));
    key += trace ? 'T' : 'S';
    if (trace && nsec > 0) {
        key.append(reinterpret_cast<const char*>(ss), sizeof(int32_t) * nsec);
        key += '|';
        key.append(reinterpret_cast<const char*>(se), sizeof(int32_t) * nsec);
    }
    for (auto& p : c->plans) {
        if (p->key == key) return p.get();
    }
    if (c->plans.size() >= 128) {      /* bounded (tuneAlignment walks 35 penalty pairs x 2 adaptors): start over once idle */
        CUDA_CHECK(cudaDeviceSynchronize());
        for (auto& p : c->plans) p->d.buf.release();
        c->plans.clear();
    }
    std::unique_ptr<sarlacc_chunk::CachedPlan> np(new sarlacc_chunk::CachedPlan());
    np->key = key;
    const char* refs[1] = {reference};
    build_plan(np->plan, c->enc, refs, 1, (int)std::strlen(reference), true, go, ge, trace);
    if (trace && nsec > 0) {
        np->plan.sec_starts.assign(ss, ss + nsec);
        np->plan.sec_ends.assign(se, se + nsec);
    }
    np->d.upload(np->plan, c->st);
    c->plans.push_back(std::move(np));
    return c->plans.back().get();
}

/* Copies a device vector to caller memory (host or device) on the chunk's copy stream. */
void chunk_copy_out(sarlacc_chunk* c, void* dst, const void* src, size_t bytes) {
    if (!dst || bytes == 0) return;
    CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, c->cp));
}

void chunk_wait_tracebacks(sarlacc_chunk* c, cudaStream_t on) {
    for (int b = 0; b < 2; ++b) {
        if (c->tb_pending[b]) CUDA_CHECK(cudaStreamWaitEvent(on, c->tb_done[b], 0));
    }
}

int chunk_check_loaded(sarlacc_chunk* c) {
    if (!c) return fail("chunk handle is NULL");
    if (c->n <= 0) return fail("the chunk holds no reads");
    return 0;
}

}  // namespace

void sarlacc_chunk_free(sarlacc_chunk* c);

sarlacc_chunk* sarlacc_chunk_create(int device, int64_t capacity, int tolerance, const sarlacc_encoding* encoding) {
    if (capacity <= 0) { fail("chunk capacity must be positive"); return nullptr; }
    if (tolerance <= 0) { fail("tolerance should be a positive integer"); return nullptr; }
    if (require_device()) return nullptr;
    std::unique_ptr<sarlacc_chunk> c(new sarlacc_chunk());
    const char* msg = build_encoding(encoding, c->enc);
    if (msg) { fail(msg); return nullptr; }
    try {
        CUDA_CHECK(cudaSetDevice(device));
        c->device = device;
        c->sms = device_sm_count(device);
        c->capacity = capacity;
        c->tol = tolerance;
        c->stride = std::max(8, (tolerance + 8) & ~7);
        CUDA_CHECK(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
        CUDA_CHECK(cudaStreamCreateWithFlags(&c->tb, cudaStreamNonBlocking));
        CUDA_CHECK(cudaStreamCreateWithFlags(&c->cp, cudaStreamNonBlocking));
        for (int k = 0; k < 2; ++k) {
            CUDA_CHECK(cudaEventCreateWithFlags(&c->fwd_done[k], cudaEventDisableTiming));
            CUDA_CHECK(cudaEventCreateWithFlags(&c->tb_done[k], cudaEventDisableTiming));
            CUDA_CHECK(cudaEventCreateWithFlags(&c->out_copied[k], cudaEventDisableTiming));
            CUDA_CHECK(cudaEventCreateWithFlags(&c->sout_copied[k], cudaEventDisableTiming));
        }
        CUDA_CHECK(cudaEventCreateWithFlags(&c->out_ready, cudaEventDisableTiming));
        const size_t cap = (size_t)capacity;
        c->rows_f.reserve(sizeof(uint16_t) * cap * c->stride);
        c->rows_b.reserve(sizeof(uint16_t) * cap * c->stride);
        c->lens_f.reserve(sizeof(int32_t) * cap);
        c->lens_b.reserve(sizeof(int32_t) * cap);
        c->width.reserve(sizeof(int32_t) * cap);
        c->flipped.reserve(cap);
        c->tmp.reserve(sizeof(double) * 4 * cap);
    } catch (CudaError& e) {
        fail(e.msg);
        sarlacc_chunk_free(c.release());
        return nullptr;
    }
    return c.release();
}

void sarlacc_chunk_free(sarlacc_chunk* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (DevBuf* b : {&c->rows_f, &c->rows_b, &c->lens_f, &c->lens_b, &c->width, &c->flipped, &c->srows_f, &c->srows_b, &c->barcodes,
                      &c->tmp, &c->out[0], &c->out[1], &c->sout[0], &c->sout[1]}) b->release();
    c->raw_f.release();
    c->raw_b.release();
    c->h_lens_f.release();
    c->h_lens_b.release();
    c->h_width.release();
    c->pair[0].release();
    c->pair[1].release();
    c->score_scratch.release();
    c->score_scratch2.release();
    c->side.release();
    for (int k = 0; k < 2; ++k) if (c->fw[k]) cudaStreamDestroy(c->fw[k]);
    if (c->rows_ready) cudaEventDestroy(c->rows_ready);
    for (auto& p : c->plans) p->d.buf.release();
    for (int k = 0; k < 2; ++k) {
        if (c->fwd_done[k]) cudaEventDestroy(c->fwd_done[k]);
        if (c->tb_done[k]) cudaEventDestroy(c->tb_done[k]);
        if (c->out_copied[k]) cudaEventDestroy(c->out_copied[k]);
        if (c->sout_copied[k]) cudaEventDestroy(c->sout_copied[k]);
    }
    if (c->out_ready) cudaEventDestroy(c->out_ready);
    for (auto e : c->tev) cudaEventDestroy(e);
    if (c->st) cudaStreamDestroy(c->st);
    if (c->tb) cudaStreamDestroy(c->tb);
    if (c->cp) cudaStreamDestroy(c->cp);
    delete c;
}

int64_t sarlacc_chunk_n(const sarlacc_chunk* c) { return c ? c->n : 0; }

int sarlacc_chunk_load_mock(sarlacc_chunk* c, int64_t n, uint64_t first_index, uint64_t seed,
        const char* adaptor1, const char* adaptor2, int insert_len, const char* const* barcodes, int nbarcodes,
        double sub_rate, double indel_rate, int max_insert)
{
    Range nvtx("sarlacc_chunk_load_mock");
    if (!c) return fail("chunk handle is NULL");
    if (n < 0 || n > c->capacity) return fail("more reads than the chunk's capacity");
    if (!adaptor1 || !adaptor2) return fail("adaptor sequence should be a string");
    const size_t l1 = std::strlen(adaptor1), l2 = std::strlen(adaptor2);
    if (l1 > 127 || l2 > 127) return fail("mock adaptors are limited to 127 bases");
    if (max_insert < 2 || max_insert > 64) return fail("max_insert out of range");
    if (!(sub_rate >= 0 && sub_rate < 1 && indel_rate >= 0 && indel_rate < 1 && sub_rate + indel_rate > 0)) return fail("rates out of range");
    try {
        CUDA_CHECK(cudaSetDevice(c->device));
        MockArgs M;
        std::memset(&M, 0, sizeof(M));
        M.n = n;
        M.seed = seed;
        M.first_index = first_index;
        M.tol = c->tol;
        M.stride = c->stride;
        M.front = c->rows_f.as<uint16_t>();
        M.back = c->rows_b.as<uint16_t>();
        M.lens_front = c->lens_f.as<int32_t>();
        M.lens_back = c->lens_b.as<int32_t>();
        M.width = c->width.as<int32_t>();
        M.flipped = c->flipped.as<uint8_t>();
        M.len1 = (int)l1;
        M.len2 = (int)l2;
        for (size_t i = 0; i < l1; ++i) M.adaptor1[i] = (char)std::toupper((unsigned char)adaptor1[i]);
        for (size_t i = 0; i < l2; ++i) M.adaptor2[i] = (char)std::toupper((unsigned char)adaptor2[i]);
        /* the barcode slot: adaptor1's first run of N (R/mockReads.R:35-50) */
        M.run0_start = M.run0_end = -1;
        for (size_t i = 0; i < l1; ++i) {
            if (M.adaptor1[i] == 'N') {
                size_t j = i;
                while (j < l1 && M.adaptor1[j] == 'N') ++j;
                M.run0_start = (int)i;
                M.run0_end = (int)j;
                break;
            }
        }
        M.nbarcodes = 0;
        if (nbarcodes > 0 && barcodes && M.run0_start >= 0) {
            const int bl = M.run0_end - M.run0_start;
            std::vector<uint8_t> codes((size_t)nbarcodes * bl);
            for (int b = 0; b < nbarcodes; ++b) {
                if (!barcodes[b] || (int)std::strlen(barcodes[b]) != bl) return fail("barcodes must have the length of adaptor1's first N run");
                for (int k = 0; k < bl; ++k) {
                    const char ch = (char)std::toupper((unsigned char)barcodes[b][k]);
                    const char* at = std::strchr("ACGT", ch);
                    if (!at || !ch) return fail("barcodes must consist of A, C, G, T");
                    codes[(size_t)b * bl + k] = (uint8_t)(at - "ACGT");
                }
            }
            CUDA_CHECK(cudaStreamSynchronize(c->st));      /* an earlier load may still read the previous table */
            c->barcodes.reserve(codes.size());
            CUDA_CHECK(cudaMemcpy(c->barcodes.p, codes.data(), codes.size(), cudaMemcpyHostToDevice));
            M.barcodes = c->barcodes.as<uint8_t>();
            M.nbarcodes = nbarcodes;
        }
        M.molecule_len = (int)(l1 + l2) + insert_len;
        M.max_insert = max_insert;
        M.sub_thr = (uint32_t)std::floor(sub_rate * 65536.0);
        M.indel_thr = (uint32_t)std::floor(indel_rate * 65536.0);
        /* quality = clamp(round(-10 log10(u * max_err)), 0, 93), u = word / 2^32 (R/mockReads.R:82):
         * quality >= k  <=>  word < 2^32 * 10^(-(k - 0.5) / 10) / max_err */
        const double max_err = sub_rate + indel_rate;
        M.qmin = 0;
        for (int k = 0; k <= 94; ++k) {
            double t = k == 0 ? 4294967296.0 : std::floor(4294967296.0 * std::pow(10.0, -((double)k - 0.5) / 10.0) / max_err);
            if (k == 94) t = 0.0;
            M.qthr[k] = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
            if (k <= 93 && t >= 4294967296.0) M.qmin = (uint32_t)k;
        }
        {   /* number of indels over the whole molecule: Binomial(molecule_len, p) with p = indel_thr / 65536, as an integer
             * distribution function around its mean.  Only IEEE multiplications, divisions and additions in a fixed order
             * (built with -ffp-contract=off), so that sarlacc_b200/synth.py: width_table gets the same integers. */
            const int nmol = M.molecule_len;
            const double pr = (double)M.indel_thr / 65536.0, qr = 1.0 - pr;
            const int mean = (int)((double)nmol * pr);
            M.wlo = std::max(0, mean - 128);
            double x = 1.0;
            for (int i = 0; i < nmol; ++i) x *= qr;              /* P(0 indels) */
            double cdf = 0.0;
            for (int k = 0; k < M.wlo + 256; ++k) {
                cdf += x;
                if (k >= M.wlo) {
                    const double t = std::floor(cdf * 4294967296.0);
                    M.wcdf[k - M.wlo] = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
                }
                const double t1 = (double)(nmol - k) * pr, t2 = (double)(k + 1) * qr;
                x = x * t1;
                x = x / t2;
            }
        }
        chunk_wait_tracebacks(c, c->st);      /* tracebacks of the previous contents still read the window lengths */
        chunk_mark(c, 0, true);
        launch_mock_windows(M, c->st);
        chunk_mark(c, 0, false);
        g_launches += 1;
        CUDA_CHECK(cudaGetLastError());
        c->n = n;
        c->scrambled_n = -1;
        c->maxlen = c->tol;
        c->has_width = true;
    } catch (CudaError& e) {
        return fail(e.msg);
    }
    return 0;
}

/* Host reads in: either pre-cut windows (tolerance == 0: `front` / `back` as .get_front_and_back made them, widths
 * optional) or whole reads (tolerance > 0, back == NULL: both windows cut by the device packer).  Synchronous up to the
 * point where the reference's per-read errors are known. */
int sarlacc_chunk_load_reads(sarlacc_chunk* c, const sarlacc_reads* front, const sarlacc_reads* back, int tolerance, const int32_t* width)
{
    Range nvtx("sarlacc_chunk_load_reads");
    if (!c) return fail("chunk handle is NULL");
    if (!front) return fail("reads must not be NULL");
    if (tolerance < 0 || tolerance > c->tol) return fail("tolerance exceeds the chunk's");
    if (tolerance == 0 && (!back || back->n != front->n)) return fail("front and back windows should have the same length");
    const int64_t n = front->n;
    if (n > c->capacity) return fail("more reads than the chunk's capacity");
    try {
        CUDA_CHECK(cudaSetDevice(c->device));
        ReadView VF{front}, VB{back ? back : front};
        if (tolerance > 0) {
            VF.tol = tolerance;
            VB = ReadView{front, tolerance, true};
        }
        PackTables PF, PB;
        build_pack_tables(PF, front->seq_encoding, c->enc);
        build_pack_tables(PB, VB.R->seq_encoding, c->enc);
        const int nthreads = host_threads_for(1);
        CUDA_CHECK(cudaStreamSynchronize(c->st));     /* the staging buffers of the previous load */
        chunk_wait_tracebacks(c, c->st);
        c->h_lens_f.reserve(sizeof(int32_t) * (size_t)std::max<int64_t>(n, 1));
        c->h_lens_b.reserve(sizeof(int32_t) * (size_t)std::max<int64_t>(n, 1));
        FirstError ef, eb;
        int maxf = 0, maxb = 0;
        scan_lengths(VF, 0, n, c->h_lens_f.as<int32_t>(), nthreads, ef, maxf);
        scan_lengths(VB, 0, n, c->h_lens_b.as<int32_t>(), nthreads, eb, maxb);
        if (std::max(maxf, maxb) > c->tol) return fail("a window is longer than the chunk's tolerance");
        if (ef.kind != ERR_NONE || eb.kind != ERR_NONE) return fail(err_text(ef.kind != ERR_NONE ? ef.kind : eb.kind));
        if (n > 0) {
            chunk_mark(c, 0, true);
            CUDA_CHECK(cudaMemcpyAsync(c->lens_f.p, c->h_lens_f.p, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, c->st));
            CUDA_CHECK(cudaMemcpyAsync(c->lens_b.p, c->h_lens_b.p, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, c->st));
            const bool pinned_f = !VF.R->seq && pointer_is_pinned(VF.R->seq_pool) && pointer_is_pinned(VF.R->qual_pool);
            const bool pinned_b = !VB.R->seq && pointer_is_pinned(VB.R->seq_pool) && pointer_is_pinned(VB.R->qual_pool);
            PendingPack pf, pb;
            stage_and_pack(VF, 0, n, PF, c->h_lens_f.as<int32_t>(), c->lens_f.as<int32_t>(), c->stride, c->rows_f.as<uint16_t>(), true,
                           pinned_f, c->raw_f, c->st, nthreads, &pf);
            stage_and_pack(VB, 0, n, PB, c->h_lens_b.as<int32_t>(), c->lens_b.as<int32_t>(), c->stride, c->rows_b.as<uint16_t>(), true,
                           pinned_b, c->raw_b, c->st, nthreads, &pb);
            launch_pending_pack(pf, c->st);
            launch_pending_pack(pb, c->st);
            c->has_width = tolerance > 0 || width != nullptr;
            if (c->has_width) {
                c->h_width.reserve(sizeof(int32_t) * (size_t)n);
                int32_t* hw = c->h_width.as<int32_t>();
                for (int64_t i = 0; i < n; ++i) hw[i] = tolerance > 0 ? (int32_t)VF.full_seq_len(i) : width[i];
                CUDA_CHECK(cudaMemcpyAsync(c->width.p, hw, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, c->st));
            }
            chunk_mark(c, 0, false);
            CUDA_CHECK(cudaStreamSynchronize(c->st));
            const long long fb = c->raw_f.first_bad(), fb2 = c->raw_b.first_bad();
            if (fb >= 0 || fb2 >= 0) {
                c->n = 0;
                return fail(err_text(ERR_QUAL));
            }
        }
        c->n = n;
        c->scrambled_n = -1;
        c->maxlen = std::max(maxf, maxb);
    } catch (CudaError& e) {
        return fail(e.msg);
    }
    return 0;
}

int sarlacc_chunk_adaptor_align(sarlacc_chunk* c, double gapopen, double gapext, const char* adaptor1, const char* adaptor2,
        int nsec1, const int32_t* sec_starts1, const int32_t* sec_ends1,
        int nsec2, const int32_t* sec_starts2, const int32_t* sec_ends2,
        int64_t out_pitch, int32_t* read_width, uint8_t* reversed,
        double* score1, int32_t* start1, int32_t* end1, int32_t* sec_start1, int32_t* sec_width1,
        double* score2, int32_t* start2, int32_t* end2, int32_t* sec_start2, int32_t* sec_width2)
{
    Range nvtx("sarlacc_chunk_adaptor_align");
    if (chunk_check_loaded(c)) return 1;
    if (!adaptor1 || !adaptor2) return fail("adaptor sequence should be a string");
    if (!*adaptor1 || !*adaptor2) return fail("chunk runs need two non-empty adaptors");
    if (nsec1 < 0 || nsec2 < 0 || (nsec1 > 0 && (!sec_starts1 || !sec_ends1)) || (nsec2 > 0 && (!sec_starts2 || !sec_ends2)))
        return fail("section starts and ends should have the same length");
    try {
        CUDA_CHECK(cudaSetDevice(c->device));
        const int64_t n = c->n;
        if (out_pitch < n) out_pitch = n;
        sarlacc_chunk::CachedPlan* cp[2] = {chunk_plan(c, adaptor1, gapopen, gapext, true, nsec1, sec_starts1, sec_ends1),
                                            chunk_plan(c, adaptor2, gapopen, gapext, true, nsec2, sec_starts2, sec_ends2)};
        for (int k = 0; k < 2; ++k) {
            const Plan& P = cp[k]->plan;
            if (P.bad_col[0] >= 0) return fail(err_text(ERR_REF));
            for (size_t x = 0; x < P.sec_starts.size(); ++x) {
                if (P.sec_starts[x] < 0 || P.sec_starts[x] > P.L || P.sec_ends[x] < 0 || P.sec_ends[x] > P.L) return fail("section bounds outside the adaptor");
            }
        }
        const Plan* plan[2] = {&cp[0]->plan, &cp[1]->plan};
        const DevPlan* D[2] = {&cp[0]->d, &cp[1]->d};
        const int nsec[2] = {nsec1, nsec2};
        /* final columns of this call: [n] vectors and [nsec][n] matrices in one block, double-buffered across calls */
        const int w = c->out_which;
        c->out_which ^= 1;
        auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
        size_t o_rev = 0, o_score[2], o_start[2], o_end[2], o_ss[2], o_sw[2], at = al((size_t)n);
        for (int k = 0; k < 2; ++k) {
            o_score[k] = at; at += al(sizeof(double) * (size_t)n);
            o_start[k] = at; at += al(sizeof(int32_t) * (size_t)n);
            o_end[k] = at; at += al(sizeof(int32_t) * (size_t)n);
            o_ss[k] = at; at += al(sizeof(int32_t) * (size_t)n * std::max(1, nsec[k]));
            o_sw[k] = at; at += al(sizeof(int32_t) * (size_t)n * std::max(1, nsec[k]));
        }
        if (c->out_pending[w]) {       /* the copy-out of two calls ago still reads this block */
            CUDA_CHECK(cudaStreamWaitEvent(c->st, c->out_copied[w], 0));
            c->out_pending[w] = false;
        }
        if (at > c->out[w].cap) CUDA_CHECK(cudaStreamSynchronize(c->cp));
        c->out[w].reserve(at);
        uint8_t* d = c->out[w].as<uint8_t>();

        /* sub-ranges: whole grid-fulls for both adaptors' kernels, four record sets within half the scratch budget each
         * parity, so that the tracebacks of sub-range k run beside the forward passes of k+1 */
        const long long g1 = plan_groups(*plan[0], true), g2 = plan_groups(*plan[1], true);
        long long sub = chunk_for(1 << 17, g1, g2);
        const long long fit = std::max<long long>(1, (long long)(scratch_budget_bytes() / pair_scratch_per_read(*plan[0], *plan[1], c->maxlen)));
        if (sub > fit) sub = std::max<long long>(1, whole_rounds(fit, std::max(g1, g2)));
        const char* ce = std::getenv("SARLACC_CHUNK");
        if (ce && std::atoll(ce) > 0) sub = std::atoll(ce);
        chunk_mark(c, 1, true);
        /* Consecutive sub-ranges use different scratch, so their forward passes need not wait for each other: each parity
         * has its own stream, and the first blocks of sub-range k+1 start in the tail of sub-range k's last launches. */
        const bool apart = overlap_launches() && n > sub;
        bool forked[2] = {false, false};
        if (apart) {
            for (int k = 0; k < 2; ++k) if (!c->fw[k]) CUDA_CHECK(cudaStreamCreateWithFlags(&c->fw[k], cudaStreamNonBlocking));
            if (!c->rows_ready) CUDA_CHECK(cudaEventCreateWithFlags(&c->rows_ready, cudaEventDisableTiming));
            CUDA_CHECK(cudaEventRecord(c->rows_ready, c->st));
        }
        const char* name = "";
        for (long long off = 0; off < n; off += sub) {
            const long long m = std::min<long long>(sub, n - off);
            const int b = c->parity;
            c->parity ^= 1;
            cudaStream_t fs = apart ? c->fw[b] : c->st;
            if (apart && !forked[b]) {
                CUDA_CHECK(cudaStreamWaitEvent(fs, c->rows_ready, 0));
                forked[b] = true;
            }
            if (c->tb_pending[b]) {
                CUDA_CHECK(cudaStreamWaitEvent(fs, c->tb_done[b], 0));
                c->tb_pending[b] = false;
            }
            PairDeviceOut po;
            po.reversed = d + o_rev + off;
            for (int k = 0; k < 2; ++k) {
                po.score[k] = reinterpret_cast<double*>(d + o_score[k]) + off;
                po.start[k] = reinterpret_cast<int32_t*>(d + o_start[k]) + off;
                po.end[k] = reinterpret_cast<int32_t*>(d + o_end[k]) + off;
                po.sec_start[k] = reinterpret_cast<int32_t*>(d + o_ss[k]) + off;
                po.sec_width[k] = reinterpret_cast<int32_t*>(d + o_sw[k]) + off;
            }
            po.pitch = n;
            name = run_pair_device(plan, D, c->pair[b], fs, c->tb, c->fwd_done[b], c->tb_done[b],
                                   c->rows_f.as<uint16_t>() + (size_t)off * c->stride, c->lens_f.as<int32_t>() + off, c->stride,
                                   c->rows_b.as<uint16_t>() + (size_t)off * c->stride, c->lens_b.as<int32_t>() + off, c->stride,
                                   m, c->maxlen, c->has_width ? c->width.as<int32_t>() + off : nullptr,
                                   c->tmp.as<double>() + 4 * (size_t)off, po, c->sms);
            c->tb_pending[b] = true;
        }
        for (int k = 0; k < 2; ++k) if (forked[k]) CUDA_CHECK(cudaStreamWaitEvent(c->st, c->fwd_done[k], 0));
        chunk_mark(c, 1, false);
        for (int k = 0; k < 2; ++k) {
            const Geometry g = geometry_for(*plan[k], c->maxlen);
            c->last_kernel[k] = std::string(k == 0 ? name : g_last_kernel) + " G=" + std::to_string(g.G) + " C=" + std::to_string(g.C);
        }
        /* copy-out on the copy stream, behind the tracebacks (which are behind the forward passes and strand resolution) */
        for (int b = 0; b < 2; ++b) {
            if (c->tb_pending[b]) CUDA_CHECK(cudaStreamWaitEvent(c->cp, c->tb_done[b], 0));
        }
        chunk_copy_out(c, reversed, d + o_rev, (size_t)n);
        double* const sc[2] = {score1, score2};
        int32_t* const stt[2] = {start1, start2};
        int32_t* const enn[2] = {end1, end2};
        int32_t* const sst[2] = {sec_start1, sec_start2};
        int32_t* const sww[2] = {sec_width1, sec_width2};
        for (int k = 0; k < 2; ++k) {
            chunk_copy_out(c, sc[k], d + o_score[k], sizeof(double) * (size_t)n);
            chunk_copy_out(c, stt[k], d + o_start[k], sizeof(int32_t) * (size_t)n);
            chunk_copy_out(c, enn[k], d + o_end[k], sizeof(int32_t) * (size_t)n);
            for (int x = 0; x < nsec[k]; ++x) {
                if (sst[k]) chunk_copy_out(c, sst[k] + (size_t)x * out_pitch, d + o_ss[k] + sizeof(int32_t) * (size_t)x * n, sizeof(int32_t) * (size_t)n);
                if (sww[k]) chunk_copy_out(c, sww[k] + (size_t)x * out_pitch, d + o_sw[k] + sizeof(int32_t) * (size_t)x * n, sizeof(int32_t) * (size_t)n);
            }
        }
        if (read_width && c->has_width) chunk_copy_out(c, read_width, c->width.p, sizeof(int32_t) * (size_t)n);
        CUDA_CHECK(cudaEventRecord(c->out_copied[w], c->cp));
        c->out_pending[w] = true;
    } catch (CudaError& e) {
        return fail(e.msg);
    }
    return 0;
}

int sarlacc_chunk_scrambled_scores(sarlacc_chunk* c, double gapopen, double gapext, const char* adaptor1, const char* adaptor2,
        uint64_t seed, uint64_t first_index, const uint64_t* read_index, int scramble, double* score1, double* score2, double* strand_score)
{
    Range nvtx("sarlacc_chunk_scrambled_scores");
    if (chunk_check_loaded(c)) return 1;
    if (!adaptor1 || !adaptor2) return fail("adaptor sequence should be a string");
    if (!*adaptor1 || !*adaptor2) return fail("chunk runs need two non-empty adaptors");
    try {
        CUDA_CHECK(cudaSetDevice(c->device));
        const int64_t n = c->n;
        sarlacc_chunk::CachedPlan* cp[2] = {chunk_plan(c, adaptor1, gapopen, gapext, false, 0, nullptr, nullptr),
                                            chunk_plan(c, adaptor2, gapopen, gapext, false, 0, nullptr, nullptr)};
        if (cp[0]->plan.bad_col[0] >= 0 || cp[1]->plan.bad_col[0] >= 0) return fail(err_text(ERR_REF));
        const uint16_t* rf = c->rows_f.as<uint16_t>();
        const uint16_t* rb = c->rows_b.as<uint16_t>();
        if (scramble == 2) {          /* the permuted windows of the previous call (tuneAlignment: one scramble, 35 penalty pairs) */
            if (!c->srows_f.p || !c->srows_b.p || c->scrambled_n != n) return fail("no scrambled windows to reuse");
            rf = c->srows_f.as<uint16_t>();
            rb = c->srows_b.as<uint16_t>();
        } else if (scramble) {
            const size_t bytes = sizeof(uint16_t) * (size_t)c->capacity * c->stride;
            c->srows_f.reserve(bytes);
            c->srows_b.reserve(bytes);
            DevBuf idx;
            if (read_index) {
                idx.reserve(sizeof(uint64_t) * (size_t)n);
                CUDA_CHECK(cudaMemcpyAsync(idx.p, read_index, sizeof(uint64_t) * (size_t)n, cudaMemcpyDefault, c->st));
            }
            chunk_mark(c, 2, true);
            launch_scramble(rf, c->srows_f.as<uint16_t>(), c->lens_f.as<int32_t>(), n, c->stride, seed, first_index,
                            read_index ? idx.as<unsigned long long>() : nullptr, 0, c->st);
            launch_scramble(rb, c->srows_b.as<uint16_t>(), c->lens_b.as<int32_t>(), n, c->stride, seed, first_index,
                            read_index ? idx.as<unsigned long long>() : nullptr, 1, c->st);
            chunk_mark(c, 2, false);
            g_launches += 2;
            CUDA_CHECK(cudaGetLastError());
            if (read_index) {
                CUDA_CHECK(cudaStreamSynchronize(c->st));
                idx.release();
            }
            rf = c->srows_f.as<uint16_t>();
            rb = c->srows_b.as<uint16_t>();
            c->scrambled_n = n;
        }
        /* the four forward passes of .get_alignment_scores (R/tuneAlignment.R:99-112): START, END, RSTART, REND */
        chunk_mark(c, 3, true);
        double* tmp = c->tmp.as<double>();
        cudaStream_t side = c->side.begin(c->st);
        for (int r = 0; r < 4; ++r) {
            const int a = (r == 0 || r == 2) ? 0 : 1;
            const bool on_front = (r == 0 || r == 3);
            Outputs dev;
            dev.score = tmp + (size_t)r * n;
            forward_once(cp[a]->plan, cp[a]->d, r >= 2 ? c->score_scratch2 : c->score_scratch, r >= 2 ? side : c->st, on_front ? rf : rb,
                         on_front ? c->lens_f.as<int32_t>() : c->lens_b.as<int32_t>(), n, c->stride, c->maxlen, false, dev, c->sms);
        }
        c->side.end(c->st);
        /* kept scores: straight into device destinations, through a chunk buffer + the copy stream for host ones */
        cudaPointerAttributes at1, at2, at3;
        const bool dev1 = score1 && cudaPointerGetAttributes(&at1, score1) == cudaSuccess && at1.type == cudaMemoryTypeDevice;
        const bool dev2 = score2 && cudaPointerGetAttributes(&at2, score2) == cudaSuccess && at2.type == cudaMemoryTypeDevice;
        const bool dev3 = strand_score && cudaPointerGetAttributes(&at3, strand_score) == cudaSuccess && at3.type == cudaMemoryTypeDevice;
        cudaGetLastError();
        const int w = c->sout_which;
        c->sout_which ^= 1;
        if (c->sout_pending[w]) {
            CUDA_CHECK(cudaStreamWaitEvent(c->st, c->sout_copied[w], 0));
            c->sout_pending[w] = false;
        }
        c->sout[w].reserve(sizeof(double) * 3 * (size_t)c->capacity);
        double* own = c->sout[w].as<double>();
        StrandArgs SA;
        std::memset(&SA, 0, sizeof(SA));
        SA.n = n;
        SA.a1_front = tmp;
        SA.a2_back = tmp + (size_t)n;
        SA.a1_back = tmp + (size_t)2 * n;
        SA.a2_front = tmp + (size_t)3 * n;
        SA.reversed = nullptr;
        SA.score1 = dev1 ? score1 : (score1 ? own : nullptr);
        SA.score2 = dev2 ? score2 : (score2 ? own + c->capacity : nullptr);
        SA.strand_score = dev3 ? strand_score : (strand_score ? own + 2 * c->capacity : nullptr);
        launch_resolve_strand(SA, c->st);
        chunk_mark(c, 3, false);
        g_launches += 1;
        CUDA_CHECK(cudaGetLastError());
        if ((score1 && !dev1) || (score2 && !dev2) || (strand_score && !dev3)) {
            CUDA_CHECK(cudaEventRecord(c->out_ready, c->st));
            CUDA_CHECK(cudaStreamWaitEvent(c->cp, c->out_ready, 0));
            if (score1 && !dev1) chunk_copy_out(c, score1, own, sizeof(double) * (size_t)n);
            if (score2 && !dev2) chunk_copy_out(c, score2, own + c->capacity, sizeof(double) * (size_t)n);
            if (strand_score && !dev3) chunk_copy_out(c, strand_score, own + 2 * c->capacity, sizeof(double) * (size_t)n);
            CUDA_CHECK(cudaEventRecord(c->sout_copied[w], c->cp));
            c->sout_pending[w] = true;
        }
    } catch (CudaError& e) {
        return fail(e.msg);
    }
    return 0;
}

/* Makes the chunk's compute stream wait for everything enqueued so far on its traceback and copy streams, so that an
 * event recorded on sarlacc_chunk_stream() afterwards marks the completion of all of it (bench.py times steps with CUDA
 * events on that stream). */
int sarlacc_chunk_join(sarlacc_chunk* c) {
    if (!c) return fail("chunk handle is NULL");
    try {
        CUDA_CHECK(cudaSetDevice(c->device));
        CUDA_CHECK(cudaEventRecord(c->out_ready, c->tb));
        CUDA_CHECK(cudaStreamWaitEvent(c->st, c->out_ready, 0));
        CUDA_CHECK(cudaEventRecord(c->out_ready, c->cp));
        CUDA_CHECK(cudaStreamWaitEvent(c->st, c->out_ready, 0));
    } catch (CudaError& e) {
        return fail(e.msg);
    }
    return 0;
}

void* sarlacc_chunk_stream(sarlacc_chunk* c) { return c ? (void*)c->st : nullptr; }

int sarlacc_chunk_sync(sarlacc_chunk* c) {
    Range nvtx("sarlacc_chunk_sync");
    if (!c) return fail("chunk handle is NULL");
    try {
        CUDA_CHECK(cudaSetDevice(c->device));
        CUDA_CHECK(cudaStreamSynchronize(c->st));
        CUDA_CHECK(cudaStreamSynchronize(c->tb));
        CUDA_CHECK(cudaStreamSynchronize(c->cp));
        for (int b = 0; b < 2; ++b) c->tb_pending[b] = c->out_pending[b] = c->sout_pending[b] = false;
        chunk_collect_timing(c);
    } catch (CudaError& e) {
        return fail(e.msg);
    }
    return 0;
}

/* Tests / reports: packed rows of window set `which` (0 front, 1 back, 2 scrambled front, 3 scrambled back), lengths,
 * widths and strand flips of the loaded reads; any pointer may be NULL. */
int sarlacc_chunk_rows(sarlacc_chunk* c, int which, uint16_t* rows, int32_t* lens, int* stride, int32_t* width, uint8_t* flipped) {
    if (!c) return fail("chunk handle is NULL");
    if (which < 0 || which > 3) return fail("window set must be 0..3");
    try {
        CUDA_CHECK(cudaSetDevice(c->device));
        if (sarlacc_chunk_sync(c)) return 1;
        if (stride) *stride = c->stride;
        const DevBuf& r = which == 0 ? c->rows_f : (which == 1 ? c->rows_b : (which == 2 ? c->srows_f : c->srows_b));
        const DevBuf& l = (which & 1) ? c->lens_b : c->lens_f;
        if (c->n > 0 && rows) {
            if (!r.p) return fail("that window set has not been produced");
            CUDA_CHECK(cudaMemcpy(rows, r.p, sizeof(uint16_t) * (size_t)c->n * c->stride, cudaMemcpyDeviceToHost));
        }
        if (c->n > 0 && lens) CUDA_CHECK(cudaMemcpy(lens, l.p, sizeof(int32_t) * (size_t)c->n, cudaMemcpyDeviceToHost));
        if (c->n > 0 && width && c->has_width) CUDA_CHECK(cudaMemcpy(width, c->width.p, sizeof(int32_t) * (size_t)c->n, cudaMemcpyDeviceToHost));
        if (c->n > 0 && flipped) CUDA_CHECK(cudaMemcpy(flipped, c->flipped.p, (size_t)c->n, cudaMemcpyDeviceToHost));
    } catch (CudaError& e) {
        return fail(e.msg);
    }
    return 0;
}

void sarlacc_chunk_set_timing(sarlacc_chunk* c, int on) {
    if (!c) return;
    c->timing = on != 0;
    for (double& x : c->phase_ms) x = 0;
}

int sarlacc_chunk_phase_ms(sarlacc_chunk* c, double* ms4) {
    if (!c || !ms4) return fail("chunk handle is NULL");
    if (sarlacc_chunk_sync(c)) return 1;
    for (int k = 0; k < 4; ++k) ms4[k] = c->phase_ms[k];
    return 0;
}

const char* sarlacc_chunk_last_kernel(const sarlacc_chunk* c, int adaptor) {
    return (c && (adaptor == 0 || adaptor == 1)) ? c->last_kernel[adaptor].c_str() : "";
}

}  // extern "C"
