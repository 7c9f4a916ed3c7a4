/* CUDA kernels (sm_100a) for the quality-weighted adaptor alignment.
 *
 * Replaces the arithmetic of reference_align::align / align_column / backtrack / fill_map
 * (/root/reference/src/reference_align.cpp:54-181, 231-350).  Everything is FP64 with the
 * reference's operation order and comparison operators, compiled with --fmad=false and explicit
 * *_rn intrinsics, because tie-breaking between co-optimal paths depends on the last bit.
 *
 * Forward kernel `wf_forward<C,TRACE,ALT>` ("wavefront"):
 *   - a group of G lanes (G = 1..32, power of two, chosen per reference length) owns one alignment;
 *     lane j owns C consecutive reference columns j*C+1 .. j*C+C, state H[i-1][c], F[i-1][c] in registers;
 *   - read rows stream through the group: lane j works on row t-j at step t and hands H/E of its last
 *     column to lane j+1 with two 64-bit warp shuffles -- the anti-diagonal wavefront;
 *   - alignments are chained back to back (lane 0 starts the next alignment while lane G-1 finishes
 *     the previous one), so the pipeline never drains and short references cost no skew;
 *   - per cell it records 4 bits {pd,p5,p1,p2}; the traceback kernel turns them back into the
 *     reference's path.  See DESIGN.md for why these 4 bits are equivalent to the reference's
 *     int direction matrix with jump lengths.
 *
 * Generic kernel `generic_forward`: one thread per alignment, literal recurrence, state in global
 * memory.  Used for shapes/parameters outside the wavefront kernel's envelope (reference longer
 * than 384, more than one IUPAC class, negative gap opening ...).  Slow but exact, and still CUDA.
 */
#include "kernels.h"

#include <cstdlib>
#include <type_traits>

namespace sarlacc {

namespace {

#ifndef SARLACC_WF_BLOCK
#define SARLACC_WF_BLOCK 128
#endif
constexpr int kBlock = SARLACC_WF_BLOCK;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ double neg_inf() { return __longlong_as_double(0xfff0000000000000LL); }

/* H[i][0]: src/reference_align.cpp:65-74 */
__device__ __forceinline__ double col0_value(int local, double gop, double ge, int i) {
    if (local || i == 0) return 0.0;
    return __dsub_rn(-gop, __dmul_rn(ge, (double)(i - 1)));
}

template <int C>
struct FlagWord {   /* 4 bits per owned column: the low word holds the first 8 (32-bit) or 16 (64-bit) columns ... */
    using type = typename std::conditional<(C <= 8), uint32_t, unsigned long long>::type;
};
template <int C, bool SOLO>
struct FlagHi {     /* ... columns 16.. go to a second plane indexed alike: one byte (C = 17, 18) or one 32-bit word (solo) */
    using type = typename std::conditional<SOLO, uint32_t, uint8_t>::type;
};

#ifndef SARLACC_WF_STEP_UNROLL
#define SARLACC_WF_STEP_UNROLL 4   /* rows per trip of the inner loop: lets ptxas rename loop-carried state instead of moving it (profiles/) */
#endif
constexpr int kStepUnroll = SARLACC_WF_STEP_UNROLL;
#ifndef SARLACC_WF_PAIR_UNROLL
#define SARLACC_WF_PAIR_UNROLL 2
#endif
constexpr int kPairUnroll = SARLACC_WF_PAIR_UNROLL;
#ifndef SARLACC_WF_PAIR_UNROLL_XL
#define SARLACC_WF_PAIR_UNROLL_XL 1   /* C > 12 */
#endif
constexpr int kPairUnrollXL = SARLACC_WF_PAIR_UNROLL_XL;
#ifndef SARLACC_WF_BLOCKS_SMALL
#define SARLACC_WF_BLOCKS_SMALL 4   /* resident 128-thread blocks per SM for C <= 9 (register cap 128) */
#endif
#ifndef SARLACC_WF_BLOCKS_LARGE
#define SARLACC_WF_BLOCKS_LARGE 4   /* C >= 10: 4 blocks (cap 128 registers) measured faster than 3 (cap 168) */
#endif
#ifndef SARLACC_WF_BLOCKS_XL
#define SARLACC_WF_BLOCKS_XL 3      /* C = 14, 16, 18 (register cap 168): fewer, fatter lanes; pays off without trace records */
#endif
#ifndef SARLACC_WF_BLOCKS_SOLO
#define SARLACC_WF_BLOCKS_SOLO 3    /* C = 20..24, one thread per alignment */
#endif
template <int C>
struct WfBounds {
    static constexpr int min_blocks = (C <= 9) ? SARLACC_WF_BLOCKS_SMALL : ((C <= 12) ? SARLACC_WF_BLOCKS_LARGE : ((C <= 18) ? SARLACC_WF_BLOCKS_XL : SARLACC_WF_BLOCKS_SOLO));
};

/* R/barcodeAlign.R:28-34: strict `>` for best, then strict `>` for next best. */
__device__ __forceinline__ void update_best(double s, int id, double& best, double& next, int& bid) {
    if (s > best) {
        bid = id;
        next = best;
        best = s;
    } else if (s > next) {
        next = s;
    }
}

/* Column layout of the wavefront kernel.  G lanes x C slots >= L columns; the pad = G*C - L (< G) surplus
 * slots are slot 0 of lanes 0..pad-1 ("dummy" slots that pass their left boundary through), so that the last
 * reference column always sits in slot C-1 of lane G-1: the one column with special vertical penalties
 * (src/reference_align.cpp:120-121) then needs per-lane, not per-slot, constants. */
__host__ __device__ __forceinline__ int wf_first_col(int j, int C, int pad) {   /* 1-based DP column of lane j's first real slot */
    return j * C - (j < pad ? j : pad) + 1;
}

__host__ __device__ __forceinline__ void wf_locate(int c, int C, int pad, int* j, int* k) {   /* DP column -> (lane, slot) */
    const int short_cols = pad * (C - 1);
    if (c <= short_cols) {
        *j = (c - 1) / (C - 1);
        *k = (c - 1) % (C - 1) + 1;
    } else {
        const int c2 = c - short_cols - 1;
        *j = pad + c2 / C;
        *k = c2 % C;
    }
}

template <int C, bool SOLO>
__device__ __forceinline__ void store_flags(typename FlagWord<C>::type* lo, typename FlagHi<C, SOLO>::type* hi, long long at, const uint32_t* fw) {
    if constexpr (C <= 8) {
        lo[at] = fw[0];
    } else {
        lo[at] = ((unsigned long long)fw[1] << 32) | fw[0];
        if constexpr (C > 16) hi[at] = (typename FlagHi<C, SOLO>::type)fw[2];
    }
}

/* Where the traceback's climb up one column lands.  The walk of `traceback` below, arriving at row i of a column
 * in state F with `pending` = p, moves up again iff (p && choice != up) || choice == up, carrying p2(i); otherwise it
 * takes the cell's own diag/left move.  With LF1(i) / LF0(i) = landing row when arriving with pending true / false:
 *   LF1(i) = p2(i) ? LF1(i-1) : LF0(i-1),   LF0(i) = (choice(i) == up) ? LF1(i) : i,   LF0(0) = LF1(0) = 0.
 * The forward kernels keep this for the LAST reference column (slot C-1 of lane G-1): in local mode its vertical gaps
 * are free (src/reference_align.cpp:120-121), so the reference's backtrack<> starts with one jump over the whole
 * unaligned read tail, and LF0(len) lets the traceback kernel start where that jump lands instead of reading one
 * record per tail row. */
__device__ __forceinline__ void land_step(int& lf0, int& lf1, bool pd, bool p5, bool p2, int row) {
    const int n1 = p2 ? lf1 : lf0;
    lf1 = n1;
    lf0 = (pd || p5) ? row : n1;
}

/* The three-way choice of src/reference_align.cpp:164-174 -- diagonal only if strictly best, horizontal only if strictly
 * greater than vertical -- as value + the two recorded predicates.
 * SARLACC_WF_SHORT_CHAIN: max(m, v) does not depend on this row's left-to-right chain, so it is taken first and the
 * chain only carries max(h, mv): one compare + select less between H[i][c-1] and H[i][c].  The value is the same
 * maximum; pd = (m > v) && (m > h) is the same predicate; p5 is recorded as (m > v) || (h > mv), which equals (h > v)
 * whenever pd is false -- the only case in which the traceback (and land_step) read it: if m <= v then mv = v, and if
 * m > v but the diagonal lost, h >= m > v. */
#ifndef SARLACC_WF_SHORT_SCORE
#define SARLACC_WF_SHORT_SCORE 1   /* 1: the score-only kernels of the lane-group geometries take the short chain too (70 bp: 1215 -> 1235 GCUPS) */
#endif
#ifndef SARLACC_WF_SHORT_CHAIN
#define SARLACC_WF_SHORT_CHAIN 0   /* 1: every row-pair kernel takes the short chain (A/B builds) */
#endif
#ifndef SARLACC_WF_JIT_M
#define SARLACC_WF_JIT_M 0
#endif
/* SHORT is chosen per instantiation (wf_forward2): measured on B200 (profiles/r02_history.md), the short chain gains
 * 5 % for the score-only solo kernel (1203 -> 1261 GCUPS on the 22-bp adaptor), nothing for the four-lane geometry, and
 * costs 14-16 % wherever trace records are written (the extra live predicates are materialised through SEL). */
/* Records the predicate (a > b) as bit pattern BIT of the trace word f.
 * The plain form, `if (a > b) f |= BIT`, compiles to a predicated add on the ALU pipe -- the pipe that already executes
 * the eight FSEL of a cell and is the busiest unit of the row loop (56 %, `math_pipe_throttle` the third stall reason,
 * profiles/r02_ncu_summary_a1_trace.txt) while the FMA pipe idles (15 %).  ON_FMA writes it as a predicated integer
 * multiply-add f = BIT * one + f instead, `one` being a 1 the compiler cannot see through (a kernel argument), so that
 * ptxas keeps an IMAD: the same instruction count, on the other pipe.  The comparison inside is the one the caller's
 * select uses; ptxas merges the two DSETP. */
#ifndef SARLACC_WF_TRACE_FMA
#define SARLACC_WF_TRACE_FMA 0
#endif
__device__ __forceinline__ void trace_bit(uint32_t& f, unsigned bit, double a, double b, bool p, unsigned one) {
#if SARLACC_WF_TRACE_FMA
    (void)p;
    asm("{ .reg .pred q;\n\t"
        "setp.gt.f64 q, %1, %2;\n\t"
        "@q mad.lo.u32 %0, %3, %4, %0; }"
        : "+r"(f) : "d"(a), "d"(b), "r"(one), "r"(bit));      /* bit is a constant after unrolling: an immediate operand */
#else
    (void)a; (void)b; (void)one;
    if (p) f |= bit;
#endif
}

template <bool SHORT>
__device__ __forceinline__ double pick_move(double h, double m, double v, bool& pd, bool& p5, double* tmax = nullptr) {
    if constexpr (SHORT) {
        const bool pmv = m > v;
        const double mv = pmv ? m : v;
        const bool q = h > mv;
        pd = pmv && (m > h);
        p5 = pmv || q;
        return q ? h : mv;
    } else {
        p5 = h > v;
        const double t = p5 ? h : v;
        pd = m > t;
        if (tmax) *tmax = t;
        return pd ? m : t;
    }
}

constexpr int kCostEntries = 7;   /* A, C, G, T, two-fold code, three-fold code, N */

/* Whether the solo instantiation for C columns computes the (mis)match candidate inside the chain loop (see wf_forward2).
 * Chosen per C from the static instruction count of the row loop (tools/variant_count.sh): ptxas 12.9 runs out of
 * predicate registers in one form or the other, differently per C, and then moves predicates through general registers
 * (P2R / LOP3 / ISETP: +10 instructions per cell).  tests/test_abi.py watches the built library for that. */
#ifndef SARLACC_SOLO_JIT_MASK
#define SARLACC_SOLO_JIT_MASK 0x0   /* bit (C - 20) set: that C uses the in-loop form */
#endif
template <int C>
struct SoloJit { static constexpr bool value = C >= 20 && C <= 24 && ((SARLACC_SOLO_JIT_MASK >> (C - 20)) & 1) != 0; };

/* Per-quality cost tables in shared memory: mx[q] = {match1[q], mismatch1[q]} (ACGT reference columns), iu[q] =
 * {mismatch2[q], match3[q]} (two-fold / three-fold codes), nq[q] = match4[q] (N) -- one LDS.128 per row, plus one or two
 * more loads if the reference has IUPAC codes.  Lanes whose rows have the same quality read the same address (a
 * broadcast); different qualities spread over 8 (16-byte entries) / 16 (8-byte entries) bank groups.  A single 64-byte
 * record per quality put every load of a warp on two bank groups: 23 % of the kernel's shared-memory wavefronts were
 * conflict replays (profiles/r02_ncu_summary.txt), and at 12 resident warps per SM the row loop already keeps the
 * shared-memory pipe 55 % busy with the per-cell cost loads.  cost = AlignArgs::cost, [5][encn]. */
struct QTables {
    const double2* mx;
    const double2* iu;
    const double* nq;
};

__device__ __forceinline__ void fill_q_tables(double2* mx, double2* iu, double* nq, const double* cost, int encn) {
    for (int q = threadIdx.x; q < encn; q += blockDim.x) {
        mx[q] = make_double2(cost[q], cost[encn + q]);
        iu[q] = make_double2(cost[2 * encn + q], cost[3 * encn + q]);
        nq[q] = cost[4 * encn + q];
    }
}

/* The lane's private table of one row's possible costs (entry e at tab[e * 32]): reference A,C,G,T -> mismatch1[q],
 * except the observed base's own entry -> match1[q] (src/reference_align.cpp:186-187); two-fold / three-fold codes and N
 * -> mismatch2 / match3 / match4 [q] whatever was observed (:188-209).  rw = quality index | one-hot base << 8. */
__device__ __forceinline__ void fill_cost_table(double* tab, const QTables& Q, unsigned rw, int kinds) {
    const unsigned q = rw & 0xffu;
    const int o = __ffs(rw >> 8);                   /* one-hot A=1,C=2,G=4,T=8 -> 1..4; 0 for anything else */
    const double2 mx = Q.mx[q];
    tab[0 * 32] = mx.y;
    tab[1 * 32] = mx.y;
    tab[2 * 32] = mx.y;
    tab[3 * 32] = mx.y;
    if (o) tab[(o - 1) * 32] = mx.x;
    if (kinds & 6) {
        const double2 iu = Q.iu[q];
        tab[4 * 32] = iu.x;
        tab[5 * 32] = iu.y;
    }
    if (kinds & 8) tab[6 * 32] = Q.nq[q];
}

template <int C, bool TRACE>
__global__ void __launch_bounds__(kBlock, WfBounds<C>::min_blocks) wf_forward(const __grid_constant__ AlignArgs A)
{
    using WT = typename FlagWord<C>::type;
    extern __shared__ double smem_d[];
    const int L = A.L, nref = A.nref, encn = A.enc_n;
    double* row0s = smem_d;                                   /* [L+1]                         */
    double* lanetab = row0s + (L + 1);                        /* [warps][kCostEntries][32] lane-private cost entries */
    double2* qmx = reinterpret_cast<double2*>(smem_d + (((size_t)(L + 1) + (kBlock / 32) * kCostEntries * 32 + 1) & ~(size_t)1));   /* see QTables */
    double2* qiu = qmx + encn;
    double* qn = reinterpret_cast<double*>(qiu + encn);
    const QTables Q{qmx, qiu, qn};
    uint8_t* refm = reinterpret_cast<uint8_t*>(qn + encn);   /* [nref][L] */
    uint8_t* refk = refm + (size_t)nref * L;                                                    /* [nref][L] */

    for (int x = threadIdx.x; x <= L; x += blockDim.x) row0s[x] = A.row0[x];
    for (int x = threadIdx.x; x < nref * L; x += blockDim.x) {
        refm[x] = A.refmask[x];
        refk[x] = A.refkind[x];
    }
    fill_q_tables(qmx, qiu, qn, A.cost, encn);
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int G = A.G;
    const int j = lane & (G - 1);
    const int gpw = 32 / G;
    const long long warp_global = (long long)blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
    const long long gidx = warp_global * gpw + lane / G;
    const long long NG = (long long)gridDim.x * (kBlock / 32) * gpw;
    const int pad = G * C - L;
    const bool skip0 = j < pad;                     /* slot 0 of this lane is a dummy */
    const int cfirst = wf_first_col(j, C, pad);     /* DP column of the first real slot */
    const double gop = A.gop, ge = A.ge;
    const int local = A.local;
    const int kinds = A.kinds;
    const double NEG = neg_inf();
    const bool first_lane = (j == 0);
    /* vertical penalties of slot C-1: zero for the last reference column in local mode (:120-121) */
    const double vo_last = (local && j == G - 1) ? 0.0 : gop;
    const double ve_last = (local && j == G - 1) ? 0.0 : ge;
    /* column 0, rows i >= 1 (src/reference_align.cpp:65-74): 0 in local mode, -gap_open - gap_ext * (i - 1) otherwise --
     * one expression for both (0 - 0 * x == 0), so the row loop has no branch on the mode */
    const double c0base = local ? 0.0 : -gop, c0step = local ? 0.0 : ge;

    /* This lane's private table of the current row's possible costs: entry e at mytab[e*32].  Slot k reads
     * the entry of its reference base: ACGT -> (obs == base ? match1 : mismatch1)[q], IUPAC classes ->
     * mismatch2 / match3 / match4 [q] regardless of obs (src/reference_align.cpp:184-212). */
    double* mytab = lanetab + (threadIdx.x >> 5) * kCostEntries * 32 + lane;
    const double* slotp[C];
    auto load_slots = [&](int b) {
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const int c = cfirst + k - (skip0 ? 1 : 0);   /* DP column of slot k (dummy slot 0 aliases column cfirst-1) */
            int e = 0;
            if (c >= 1 && c <= L) {
                const unsigned kind = refk[(size_t)b * L + c - 1];
                const unsigned mask = refm[(size_t)b * L + c - 1];
                e = (kind == COL_ACGT) ? (mask == 1 ? 0 : (mask == 2 ? 1 : (mask == 4 ? 2 : 3))) : 3 + (int)kind;
                if (kind == COL_ACGT && mask == 0) e = 0;   /* unrecognized base: host raises before results are used */
            }
            slotp[k] = mytab + e * 32;
        }
    };
    load_slots(0);

    double S[C], F[C];
#pragma unroll
    for (int k = 0; k < C; ++k) { S[k] = 0.0; F[k] = NEG; }
    double outS = 0.0, outE = NEG, diag0 = 0.0;
    double outS2 = 0.0, outE2 = NEG;   /* kSkew == 2: what this lane produced one row earlier */
    long long a = gidx - NG;     /* position in the processing order of the launch */
    long long r = 0;             /* the read it stands for (AlignArgs::index) */
    int b = nref - 1;
    int i = 0, len = 0, delay = kSkew * j;
    bool done = false;
    const uint16_t* rowp = A.rows;
    WT* const flo = reinterpret_cast<WT*>(A.flags);
    typename FlagHi<C, false>::type* const fhi = reinterpret_cast<typename FlagHi<C, false>::type*>(A.flags_hi);
    long long fidx = 0;     /* record word of the current row for this lane */
    double best = NEG, nextb = NEG;
    int bid = 0;
    int lf0 = 0, lf1 = 0;   /* last slot: where the traceback's climb from this row lands (see land_step) */

    /* One DP row of this lane's C columns.  `live` gates the only side effect (the trace store). */
    auto row_step = [&](double Sl, double El, bool live) {
        fill_cost_table(mytab, Q, live ? *rowp : 0u, kinds);   /* idle lanes (start-up, finished) take entry 0, not whatever their row pointer last saw */
        {   /* column 0 feeds the first lane: src/reference_align.cpp:64-78 (0 in local mode; uniform branch otherwise) */
            const double c0v = __dsub_rn(c0base, __dmul_rn(c0step, (double)(i - 1)));   /* H[i][0], i >= 1 */
            Sl = first_lane ? c0v : Sl;
            El = first_lane ? NEG : El;
        }
        const double Sl_in = Sl, El_in = El;
        uint32_t fw[(C + 7) / 8];
#pragma unroll
        for (int x = 0; x < (C + 7) / 8; ++x) fw[x] = 0;
        /* Phase 1 (independent of this row's left-to-right chain): for every owned column the vertical
         * candidate v = max(F[i-1][c]-ve, H[i-1][c]-vo) (:145-155) and the (mis)match candidate
         * m = H[i-1][c-1] + cost (:159).  F is updated in place, m kept for phase 2. */
        double m[C];
        bool p2last = false;
        {
            double diag = diag0;
#pragma unroll
            for (int k = 0; k < C; ++k) {
                const double vO = __dsub_rn(S[k], (k == C - 1) ? vo_last : gop);
                const double Fe = __dsub_rn(F[k], (k == C - 1) ? ve_last : ge);
                const bool p2 = Fe > vO;
                F[k] = p2 ? Fe : vO;
                m[k] = __dadd_rn(diag, *slotp[k]);
                diag = (k == 0 && skip0) ? diag0 : S[k];
                if (k == C - 1) p2last = p2;
                if (TRACE) {
                    uint32_t& f = fw[k >> 3];
                    if (p2) f |= 8u << (4 * (k & 7));
                }
            }
        }
        diag0 = Sl_in;
        /* Phase 2 (the serial chain along the row): horizontal candidate h = max(E[i][c-1]-ge, H[i][c-1]-go')
         * (:129-140; when the left cell itself chose "left", H == E bitwise and go' >= ge makes the max the
         * reference's value), then the choice (:164-174): diag only if strictly best, horizontal only if
         * strictly greater than vertical. */
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const double hO = __dsub_rn(Sl, gop);
            const double Ee = __dsub_rn(El, ge);
            const bool p1 = Ee > hO;
            const double h = p1 ? Ee : hO;
            const double v = F[k];
            const bool p5 = h > v;
            const double t = p5 ? h : v;
            const bool pd = m[k] > t;
            const double Sn = pd ? m[k] : t;
            S[k] = Sn;
            Sl = Sn;
            El = h;
            if (k == 0) {   /* a dummy slot hands its left boundary on unchanged */
                Sl = skip0 ? Sl_in : Sl;
                El = skip0 ? El_in : El;
            }
            if (TRACE) {
                uint32_t& f = fw[k >> 3];
                const int sh = 4 * (k & 7);
                if (pd) f |= 1u << sh;
                if (p5) f |= 2u << sh;
                if (p1) f |= 4u << sh;
                if (k == C - 1) land_step(lf0, lf1, pd, p5, p2last, i);
            }
        }
        outS = Sl;
        outE = El;
        if (TRACE) {
            if (live) {
                store_flags<C, false>(flo, fhi, fidx, fw);
            }
        }
    };

    while (__any_sync(FULL, !done)) {
        /* ---- bookkeeping (no DP work): pipeline start-up and switches to the next chained alignment ---- */
        bool act = !done;
        if (delay > 0) { --delay; act = false; }
        if (act && i == len) {
            /* next (alignment, reference) item of this group's chain */
            ++b;
            if (b == nref) {
                b = 0;
                a += NG;
                while (a < A.n) {                                    /* empty reads: launch_fill_empty */
                    r = A.index ? A.index[a] : a;
                    if ((len = A.lens[r]) != 0) break;
                    a += NG;
                }
            }
            if (a >= A.n) {
                done = true;
                act = false;
            } else {
                i = 0;
                rowp = A.rows + r * (long long)A.stride - 1;                                   /* advanced to row i before use */
                if (TRACE) fidx = a * A.fstride + (long long)j * (kSkew * G + 1);   /* word (i + kSkew * j) * G + j */
#pragma unroll
                for (int k = 0; k < C; ++k) {
                    const int c = cfirst + k - (skip0 ? 1 : 0);
                    S[k] = (c >= 0 && c <= L) ? row0s[c] : 0.0;   /* H[0][c] */
                    F[k] = NEG;                                   /* up_jump_score, :122 */
                }
                diag0 = row0s[cfirst - 1];                        /* H[0][cfirst-1] */
                lf0 = 0;
                lf1 = 0;
                if (nref > 1) load_slots(b);
                if (b == 0) { best = NEG; nextb = NEG; bid = 0; }
            }
        }

        /* ---- DP rows: as many as every lane of the warp can take before its next event.  An active lane's
         * next event is the end of its alignment; a lane still in start-up needs bookkeeping again after one
         * step; finished lanes do not constrain (they idle on a valid row with the trace store gated off). */
        int room;
        if (act) room = len - i;
        else room = done ? 0x7fffffff : 1;
        const int steps = __reduce_min_sync(FULL, room);
        if (steps == 0x7fffffff) break;
        const int inc = act ? 1 : 0;
        const long long finc = act ? G : 0;
        const uint16_t* keep_rowp = rowp;
        if (!act) rowp = A.rows;
#pragma unroll kStepUnroll
        for (int s = 0; s < steps; ++s) {
            /* Left boundary of this row: what lane j-1 produced one step ago. */
            /* With kSkew == 2 lane j lags lane j-1 by two rows, so the boundary of the row it is about to process was
             * produced two steps ago: the serial chains of consecutive rows of one lane then do not depend on each
             * other through the neighbour and ptxas can interleave them (two chains in flight per warp). */
            const double Sl = __shfl_up_sync(FULL, kSkew == 2 ? outS2 : outS, 1, G);
            const double El = __shfl_up_sync(FULL, kSkew == 2 ? outE2 : outE, 1, G);
            if (kSkew == 2) { outS2 = outS; outE2 = outE; }
            i += inc;
            rowp += inc;
            if (TRACE) fidx += finc;
            row_step(Sl, El, act);
        }
        if (!act) rowp = keep_rowp;

        if (act && i == len && j == G - 1) {
            const double s = S[C - 1];
            if (A.score) A.score[(long long)b * A.n + r] = s;
            if (TRACE) { if (A.endrow) A.endrow[a] = lf0; }
            if (A.best_id) {
                update_best(s, b + 1, best, nextb, bid);
                if (b == nref - 1) {
                    A.best_id[r] = bid;
                    A.best[r] = best;
                    A.next_best[r] = nextb;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * wf_forward2: the same wavefront, but a lane takes TWO read rows per step and lags its left neighbour by one such
 * step.  Cell (r+1, k-1) depends on row r only through cells the lane has already finished, so the serial chains of
 * the two rows are independent and are written interleaved: cell A(k) of row r next to cell B(k-1) of row r+1.  That
 * gives every warp two dependency chains in flight (the single-row kernel's top stall is `wait`, the fixed-latency
 * dependency of its one chain, profiles/).  An odd last row runs the same code with row B masked off.
 */
/* Dynamic distribution: the next free position of the launch's processing order (skipping empty reads, which have no DP:
 * launch_fill_empty), its read and window length.  Lanes of a warp that get here together share one atomic.  Kept out of
 * line: inlined, its temporaries cost the row loops of some instantiations a few register moves per cell. */
__device__ __noinline__ long long fetch_position(const AlignArgs& A, int lane, long long& r, int& len) {
    const long long a_begin = A.range ? (long long)A.range[0] : 0;
    const long long a_end = A.range ? (long long)A.range[1] : A.n;
    const unsigned peers = __activemask();
    const int leader = __ffs(peers) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(A.next, (unsigned long long)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    long long a = a_begin + (long long)base + __popc(peers & ((1u << lane) - 1u));
    while (a < a_end) {
        r = A.index ? A.index[a] : a;
        if ((len = A.lens[r]) != 0) break;
        a = a_begin + (long long)atomicAdd(A.next, 1ULL);
    }
    return a;
}

template <int C, bool TRACE, bool SOLO>
__global__ void __launch_bounds__(kBlock, WfBounds<C>::min_blocks) wf_forward2(const __grid_constant__ AlignArgs A)
{
    /* SOLO: one thread owns one alignment and all C = L columns (G = 1): no shuffles, no first-lane or dummy-slot
     * selects, column 0 is a constant boundary -- the geometry for 20-24 bp references (adaptor2, barcodes). */
    using WT = typename FlagWord<C>::type;
    static_assert(kSkew == 2, "the row-pair kernel lags its left neighbour by one two-row step: record layout and traceback assume SARLACC_WF_SKEW == 2");
    constexpr bool SHORT = (!TRACE && (SOLO || SARLACC_WF_SHORT_SCORE != 0)) || (SARLACC_WF_SHORT_CHAIN != 0);     /* see pick_move */
    const unsigned one = A.one;
    extern __shared__ double smem_d[];
    const int L = A.L, nref = A.nref, encn = A.enc_n;
    double* row0s = smem_d;                                   /* [L+1]                         */
    double* lanetab = row0s + (L + 1);                        /* [warps][2 rows][kCostEntries][32] lane-private cost entries */
    double2* qmx = reinterpret_cast<double2*>(smem_d + (((size_t)(L + 1) + (kBlock / 32) * 2 * kCostEntries * 32 + 1) & ~(size_t)1));   /* see QTables */
    double2* qiu = qmx + encn;
    double* qn = reinterpret_cast<double*>(qiu + encn);
    const QTables Q{qmx, qiu, qn};
    uint8_t* refm = reinterpret_cast<uint8_t*>(qn + encn);   /* [nref][L] */
    uint8_t* refk = refm + (size_t)nref * L;                                   /* [nref][L] */

    if (A.next && A.range && A.range[0] >= A.range[1]) return;      /* an empty part of a list (the usual case for re-runs) */
    for (int x = threadIdx.x; x <= L; x += blockDim.x) row0s[x] = A.row0[x];
    for (int x = threadIdx.x; x < nref * L; x += blockDim.x) {
        refm[x] = A.refmask[x];
        refk[x] = A.refkind[x];
    }
    fill_q_tables(qmx, qiu, qn, A.cost, encn);
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int G = SOLO ? 1 : A.G;
    const int j = SOLO ? 0 : (lane & (G - 1));
    const int gpw = 32 / G;
    const long long warp_global = (long long)blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
    const long long gidx = warp_global * gpw + lane / G;
    const long long NG = (long long)gridDim.x * (kBlock / 32) * gpw;
    const int pad = SOLO ? 0 : (G * C - L);
    const bool skip0 = SOLO ? false : (j < pad);
    const int cfirst = SOLO ? 1 : wf_first_col(j, C, pad);
    const double gop = A.gop, ge = A.ge;
    const int local = A.local;
    const int kinds = A.kinds;
    const double NEG = neg_inf();
    const bool first_lane = SOLO ? true : (j == 0);
    const double vo_last = (local && j == G - 1) ? 0.0 : gop;
    const double ve_last = (local && j == G - 1) ? 0.0 : ge;
    /* column 0, rows i >= 1 (src/reference_align.cpp:65-74): 0 in local mode, -gap_open - gap_ext * (i - 1) otherwise --
     * one expression for both (0 - 0 * x == 0), so the row loop has no branch on the mode */
    const double c0base = local ? 0.0 : -gop, c0step = local ? 0.0 : ge;
    const int rstep = SOLO ? 32 : G;          /* record words between consecutive rows of one lane */

    constexpr int kTab = kCostEntries * 32;                   /* doubles between the tables of row A and row B */
    double* mytab = lanetab + (threadIdx.x >> 5) * 2 * kTab + lane;
    const double* slotp[C];
    auto load_slots = [&](int b) {
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const int c = cfirst + k - (skip0 ? 1 : 0);
            int e = 0;
            if (c >= 1 && c <= L) {
                const unsigned kind = refk[(size_t)b * L + c - 1];
                const unsigned mask = refm[(size_t)b * L + c - 1];
                e = (kind == COL_ACGT) ? (mask == 1 ? 0 : (mask == 2 ? 1 : (mask == 4 ? 2 : 3))) : 3 + (int)kind;
                if (kind == COL_ACGT && mask == 0) e = 0;
            }
            slotp[k] = mytab + e * 32;
        }
    };
    load_slots(0);

    double S[C], F[C];
#pragma unroll
    for (int k = 0; k < C; ++k) { S[k] = 0.0; F[k] = NEG; }
    double outSA = 0.0, outEA = NEG, outSB = 0.0, outEB = NEG, diag0 = 0.0;
    /* Static distribution: group g takes positions g, g + NG, ... of the processing order.  Dynamic (A.next): the group's
     * first lane takes the next free position from a device counter when it starts an alignment, and lane j, which runs one
     * step behind lane j-1, takes over lane j-1's position when its own alignment ends. */
    /* (the bounds are re-read where they are needed -- bookkeeping only -- instead of being held in registers across the row loop) */
    auto first_pos = [&]() -> long long { return (A.next && A.range) ? (long long)A.range[0] : 0; };
    auto end_pos = [&]() -> long long { return (A.next && A.range) ? (long long)A.range[1] : A.n; };
    long long a = A.next ? end_pos() : gidx - NG;     /* position in the processing order of the launch */
    long long r = 0;                           /* the read it stands for (AlignArgs::index) */
    int b = nref - 1;
    int i = 0, len = 0, delay = SOLO ? 0 : j;
    bool done = false;
    const uint16_t* rowp = A.rows;          /* points at the next unprocessed row */
    unsigned cur = 0;                        /* rows i+1, i+2 (two 16-bit entries), loaded one step ahead */
    /* the next step's two rows, requested before this step's arithmetic; i stays even until the last row, so the
     * 32-bit load is aligned, and stride >= len + 1 keeps an odd last row's partner inside the window's slot */
    auto prefetch = [&](bool act_) -> unsigned {
        return (act_ && i + 2 < len) ? *reinterpret_cast<const unsigned*>(rowp + 2) : 0u;
    };
    WT* const flo = reinterpret_cast<WT*>(A.flags);
    typename FlagHi<C, SOLO>::type* const fhi = reinterpret_cast<typename FlagHi<C, SOLO>::type*>(A.flags_hi);
    long long fidx = 0;      /* record word of the next unprocessed row for this lane */
    double best = NEG, nextb = NEG;
    int bid = 0;
    int lf0 = 0, lf1 = 0;   /* last slot: landing row of the traceback's climb (land_step) */

    /* Rows i+1 (A) and i+2 (B) of this lane's C columns.  MASKED: row B may not exist (hasB false) and then must leave no
     * trace in the state.  `live` gates the trace stores. */
    auto pair_step = [&](auto masked_tag, unsigned rw2, double SlA, double ElA, double SlB, double ElB, bool live, bool hasB) {
        constexpr bool MASKED = decltype(masked_tag)::value;
        fill_cost_table(mytab, Q, rw2 & 0xffffu, kinds);
        fill_cost_table(mytab + kTab, Q, rw2 >> 16, kinds);
        {   /* column 0 feeds the first lane: src/reference_align.cpp:64-78 */
            const double c0a = __dsub_rn(c0base, __dmul_rn(c0step, (double)i));           /* H[i+1][0] */
            const double c0b = __dsub_rn(c0base, __dmul_rn(c0step, (double)(i + 1)));     /* H[i+2][0] */
            SlA = first_lane ? c0a : SlA;
            ElA = first_lane ? NEG : ElA;
            SlB = first_lane ? c0b : SlB;
            ElB = first_lane ? NEG : ElB;
        }
        const double SlA_in = SlA, ElA_in = ElA, SlB_in = SlB, ElB_in = ElB;
        uint32_t fa[(C + 7) / 8], fb[(C + 7) / 8];
#pragma unroll
        for (int x = 0; x < (C + 7) / 8; ++x) { fa[x] = 0; fb[x] = 0; }
        /* row A, phase 1: vertical and (mis)match candidates of every column (:145-159); F updated in place.
         * SARLACC_WF_JIT_M: computed per column inside the chain loop instead (no m[] array: fewer live registers). */
        /* JIT: m is computed per column inside the chain loop instead (no m[] array).  Chosen per instantiation by what
         * ptxas 12.9 makes of it: with the other form it runs out of predicate registers in the row loop of solo C = 20..22
         * (and with this form in C = 23, 24) and spills them through P2R / LOP3 / ISETP, +10 instructions per cell
         * (tests/test_abi.py::test_no_predicate_spills_in_row_loops watches the built library for that). */
        constexpr bool JIT = (SOLO && SoloJit<C>::value) || (SARLACC_WF_JIT_M != 0);
        double mA[JIT ? 1 : C];
        bool p2lastA = false;
        double diagA = diag0;
        if constexpr (!JIT) {
            double diag = diag0;
#pragma unroll
            for (int k = 0; k < C; ++k) {
                const double vO = __dsub_rn(S[k], (k == C - 1) ? vo_last : gop);
                const double Fe = __dsub_rn(F[k], (k == C - 1) ? ve_last : ge);
                const bool p2 = Fe > vO;
                F[k] = p2 ? Fe : vO;
                mA[k] = __dadd_rn(diag, *slotp[k]);
                diag = (k == 0 && skip0) ? diag0 : S[k];
                if (k == C - 1) p2lastA = p2;
                if (TRACE) trace_bit(fa[k >> 3], 8u << (4 * (k & 7)), Fe, vO, p2, one);
            }
        }
        /* the two serial chains, interleaved: A(k) and B(k-1) */
        double dB = SlA_in;               /* H[A][column left of slot 0]: diagonal of B's slot 0 */
#pragma unroll
        for (int k = 0; k <= C; ++k) {
            if (k < C) {
                double mAk;
                if constexpr (JIT) {
                    const double vO = __dsub_rn(S[k], (k == C - 1) ? vo_last : gop);
                    const double Fe = __dsub_rn(F[k], (k == C - 1) ? ve_last : ge);
                    const bool p2A = Fe > vO;
                    F[k] = p2A ? Fe : vO;
                    mAk = __dadd_rn(diagA, *slotp[k]);
                    diagA = (k == 0 && skip0) ? diag0 : S[k];
                    if (k == C - 1) p2lastA = p2A;
                    if (TRACE) trace_bit(fa[k >> 3], 8u << (4 * (k & 7)), Fe, vO, p2A, one);
                } else {
                    mAk = mA[k];
                }
                const double hO = __dsub_rn(SlA, gop);
                const double Ee = __dsub_rn(ElA, ge);
                const bool p1 = Ee > hO;
                const double h = p1 ? Ee : hO;
                bool pd, p5;
                double tA = 0.0;
                const double Sn = pick_move<SHORT>(h, mAk, F[k], pd, p5, &tA);
                S[k] = Sn;
                SlA = Sn;
                ElA = h;
                if (k == 0) {
                    SlA = skip0 ? SlA_in : SlA;
                    ElA = skip0 ? ElA_in : ElA;
                }
                if (TRACE) {
                    const int sh = 4 * (k & 7);
                    trace_bit(fa[k >> 3], 1u << sh, mAk, tA, pd, one);
                    trace_bit(fa[k >> 3], 2u << sh, h, F[k], p5, one);
                    trace_bit(fa[k >> 3], 4u << sh, Ee, hO, p1, one);
                    if (k == C - 1) land_step(lf0, lf1, pd, p5, p2lastA, i + 1);
                }
            }
            if (k >= 1) {
                constexpr int dummy = 0;
                (void)dummy;
                const int kk = k - 1;
                const double SA = S[kk], FA = F[kk];          /* row A's results in this column */
                const double vO = __dsub_rn(SA, (kk == C - 1) ? vo_last : gop);
                const double Fe = __dsub_rn(FA, (kk == C - 1) ? ve_last : ge);
                const bool p2 = Fe > vO;
                const double vB = p2 ? Fe : vO;
                const double mB = __dadd_rn(dB, slotp[kk][kTab]);
                dB = (kk == 0 && skip0) ? dB : SA;
                const double hO = __dsub_rn(SlB, gop);
                const double Ee = __dsub_rn(ElB, ge);
                const bool p1 = Ee > hO;
                const double h = p1 ? Ee : hO;
                bool pd, p5;
                double tB = 0.0;
                const double Sn = pick_move<SHORT>(h, mB, vB, pd, p5, &tB);
                if (MASKED) {
                    S[kk] = hasB ? Sn : SA;
                    F[kk] = hasB ? vB : FA;
                } else {
                    S[kk] = Sn;
                    F[kk] = vB;
                }
                SlB = Sn;
                ElB = h;
                if (kk == 0) {
                    SlB = skip0 ? SlB_in : SlB;
                    ElB = skip0 ? ElB_in : ElB;
                }
                if (TRACE) {
                    const int sh = 4 * (kk & 7);
                    trace_bit(fb[kk >> 3], 8u << sh, Fe, vO, p2, one);
                    trace_bit(fb[kk >> 3], 1u << sh, mB, tB, pd, one);
                    trace_bit(fb[kk >> 3], 2u << sh, h, vB, p5, one);
                    trace_bit(fb[kk >> 3], 4u << sh, Ee, hO, p1, one);
                    if (kk == C - 1) {
                        if (MASKED) {
                            int t0 = lf0, t1 = lf1;
                            land_step(t0, t1, pd, p5, p2, i + 2);
                            lf0 = hasB ? t0 : lf0;
                            lf1 = hasB ? t1 : lf1;
                        } else {
                            land_step(lf0, lf1, pd, p5, p2, i + 2);
                        }
                    }
                }
            }
        }
        outSA = SlA;
        outEA = ElA;
        outSB = SlB;
        outEB = ElB;
        if (MASKED) diag0 = hasB ? SlB_in : SlA_in;
        else diag0 = SlB_in;
        if (TRACE) {
            if (live) {
                store_flags<C, SOLO>(flo, fhi, fidx, fa);
                if (!MASKED || hasB) store_flags<C, SOLO>(flo, fhi, fidx + rstep, fb);
            }
        }
    };

    while (__any_sync(FULL, !done)) {
        /* ---- bookkeeping: pipeline start-up and switches to the next chained alignment ---- */
        bool act = !done;
        if (delay > 0) { --delay; act = false; }
        long long a_left = a;
        const bool dyn = A.next != nullptr;
        if (!SOLO) { if (dyn) a_left = __shfl_up_sync(FULL, a, 1, G); }      /* the left lane's position before this pass's switches */
        const long long a_end = end_pos();
        if (act && i == len) {
            ++b;
            if (b == nref) {
                b = 0;
                if (!dyn) {
                    a += NG;
                    while (a < a_end) {
                        r = A.index ? A.index[a] : a;
                        if ((len = A.lens[r]) != 0) break;
                        a += NG;
                    }
                } else if (SOLO || first_lane) {
                    a = fetch_position(A, lane, r, len);
                } else {
                    a = a_left;                       /* never an empty read: the first lane skipped those */
                    if (a < a_end) {
                        r = A.index ? A.index[a] : a;
                        len = A.lens[r];
                    }
                }
            }
            if (a >= a_end) {
                done = true;
                act = false;
            } else {
                i = 0;
                rowp = A.rows + r * (long long)A.stride;
                cur = *reinterpret_cast<const unsigned*>(rowp);
                if (TRACE) {
                    if (SOLO) fidx = ((a >> 5) * A.fstride + 1) * 32 + (a & 31);     /* [32 alignments][row][lane], row = 1 */
                    else fidx = a * A.fstride + (long long)(1 + 2 * j) * G + j;      /* word (i + 2j) * G + j, i = 1 */
                }
#pragma unroll
                for (int k = 0; k < C; ++k) {
                    const int c = cfirst + k - (skip0 ? 1 : 0);
                    S[k] = (c >= 0 && c <= L) ? row0s[c] : 0.0;
                    F[k] = NEG;
                }
                diag0 = row0s[cfirst - 1];
                lf0 = 0;
                lf1 = 0;
                if (nref > 1) load_slots(b);
                if (b == 0) { best = NEG; nextb = NEG; bid = 0; }
            }
        }

        /* ---- DP rows, two per step.  Unmasked steps run while every lane of the warp still has a whole pair of rows
         * before its next event; otherwise one masked step (odd last row, start-up) is taken. ---- */
        int room;
        if (act) room = (len - i) >> 1;
        else room = done ? 0x7fffffff : 0;
        const int steps = __reduce_min_sync(FULL, room);
        if (steps == 0x7fffffff) break;
        if (steps > 0) {
            const int inc = act ? 2 : 0;
            const long long finc = act ? 2 * rstep : 0;
#pragma unroll (C > 12 ? kPairUnrollXL : kPairUnroll)
            for (int s = 0; s < steps; ++s) {
                double SlA = 0.0, ElA = 0.0, SlB = 0.0, ElB = 0.0;     /* SOLO: column 0 is set inside pair_step */
                if (!SOLO) {
                    SlA = __shfl_up_sync(FULL, outSA, 1, G);
                    ElA = __shfl_up_sync(FULL, outEA, 1, G);
                    SlB = __shfl_up_sync(FULL, outSB, 1, G);
                    ElB = __shfl_up_sync(FULL, outEB, 1, G);
                }
                const unsigned nxt = prefetch(act);
                pair_step(std::false_type(), cur, SlA, ElA, SlB, ElB, act, true);
                cur = nxt;
                i += inc;
                rowp += inc;
                if (TRACE) fidx += finc;
            }
        } else {
            double SlA = 0.0, ElA = 0.0, SlB = 0.0, ElB = 0.0;
            if (!SOLO) {
                SlA = __shfl_up_sync(FULL, outSA, 1, G);
                ElA = __shfl_up_sync(FULL, outEA, 1, G);
                SlB = __shfl_up_sync(FULL, outSB, 1, G);
                ElB = __shfl_up_sync(FULL, outEB, 1, G);
            }
            const bool hasB = act && (len - i) >= 2;
            const unsigned nxt = prefetch(act);
            pair_step(std::true_type(), act ? (hasB ? cur : (cur & 0xffffu)) : 0u, SlA, ElA, SlB, ElB, act, hasB);   /* a missing row B reads as entry 0, never as whatever follows the window */
            cur = nxt;
            const int inc = act ? (hasB ? 2 : 1) : 0;
            i += inc;
            rowp += inc;
            if (TRACE) fidx += (long long)inc * rstep;
        }

        if (act && i == len && j == G - 1) {
            const double s = S[C - 1];
            if (A.score) A.score[(long long)b * A.n + r] = s;
            if (TRACE) { if (A.endrow) A.endrow[a] = lf0; }
            if (A.best_id) {
                update_best(s, b + 1, best, nextb, bid);
                if (b == nref - 1) {
                    A.best_id[r] = bid;
                    A.best[r] = best;
                    A.next_best[r] = nextb;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Generic fallback: literal restatement of align_column, one thread per alignment.
 * gS/gE/gChoice: [maxlen+1][gthreads] (row-major over rows so that a warp's accesses coalesce).
 */
enum { CH_DIAG = 0, CH_LEFT = 1, CH_UP = 2 };

__global__ void __launch_bounds__(128) generic_forward(const AlignArgs A, const int trace)
{
    const long long T = A.gthreads;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int L = A.L, encn = A.enc_n;
    const double gop = A.gop, ge = A.ge;
    const double NEG = neg_inf();
    double* gS = A.gS + t;
    double* gE = A.gE + t;
    uint8_t* gC = A.gChoice + t;

    for (long long a = t; a < A.n; a += T) {
        const int len = A.lens[a];
        if (len == 0) continue;
        const uint16_t* rowp = A.rows + a * (long long)A.stride;
        uint8_t* fl = trace ? reinterpret_cast<uint8_t*>(A.flags) + a * A.fstride : nullptr;
        double best = NEG, nextb = NEG;
        int bid = 0;
        for (int b = 0; b < A.nref; ++b) {
            for (int i = 0; i <= len; ++i) {
                gS[i * T] = col0_value(A.local, gop, ge, i);
                gE[i * T] = NEG;
                gC[i * T] = CH_UP;          /* directions[0..len] = -1, :64 */
            }
            for (int c = 1; c <= L; ++c) {
                const bool last = A.local && c == L;
                const unsigned mask = A.refmask[(size_t)b * L + c - 1];
                const unsigned kind = A.refkind[(size_t)b * L + c - 1];
                const double vo = last ? 0.0 : gop, ve = last ? 0.0 : ge;
                double diag = gS[0];
                double sprev = __dsub_rn(diag, gC[0] == CH_LEFT ? ge : gop);   /* :117 */
                gS[0] = sprev;
                gC[0] = CH_LEFT;                                                /* :118 */
                int chprev = CH_LEFT;
                double upS = NEG;
                for (int i = 1; i <= len; ++i) {
                    const double sold = gS[i * T];
                    double horiz = __dsub_rn(sold, gC[i * T] == CH_LEFT ? ge : gop);
                    const double ls = __dsub_rn(gE[i * T], ge);
                    const bool p1 = ls > horiz;
                    if (p1) horiz = ls;
                    gE[i * T] = horiz;
                    double vert = __dsub_rn(sprev, chprev == CH_UP ? ve : vo);
                    upS = __dsub_rn(upS, ve);
                    const bool p2 = upS > vert;
                    if (p2) vert = upS;
                    upS = vert;
                    const unsigned rw = rowp[i - 1];
                    const unsigned q = rw & 0xffu;
                    double cost;
                    if (kind == COL_ACGT) {
                        cost = ((rw >> 8) & mask) ? A.cost[q] : A.cost[encn + q];
                    } else {
                        cost = A.cost[(size_t)(1 + kind) * encn + q];
                    }
                    const double m = __dadd_rn(diag, cost);
                    diag = sold;
                    const bool p5 = horiz > vert;
                    int ch;
                    double s;
                    if (m > horiz && m > vert) {
                        ch = CH_DIAG; s = m;
                    } else if (p5) {
                        ch = CH_LEFT; s = horiz;
                    } else {
                        ch = CH_UP; s = vert;
                    }
                    gS[i * T] = s;
                    gC[i * T] = (uint8_t)ch;
                    sprev = s;
                    chprev = ch;
                    if (trace) {
                        fl[(long long)(c - 1) * len + (i - 1)] =
                            (uint8_t)((ch == CH_DIAG ? 1u : 0u) | (p5 ? 2u : 0u) | (p1 ? 4u : 0u) | (p2 ? 8u : 0u));
                    }
                }
            }
            const double s = gS[(long long)len * T];
            if (A.score) A.score[(long long)b * A.n + a] = s;
            if (A.best_id) update_best(s, b + 1, best, nextb, bid);
        }
        if (A.best_id) {
            A.best_id[a] = bid;
            A.best[a] = best;
            A.next_best[a] = nextb;
        }
    }
}

/* Empty reads need no DP: the score is the row-0 chain H[0][L] (e.g. -(16+5) for a 16-mer, go=5, ge=1;
 * tests/testthat/test-adaptor-align.R:53-56). */
__global__ void fill_empty(const AlignArgs A)
{
    const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= A.n || A.lens[a] != 0) return;
    const double s = A.row0[A.L];
    for (int b = 0; b < A.nref; ++b) {
        if (A.score) A.score[(long long)b * A.n + a] = s;
    }
    if (A.best_id) {
        A.best_id[a] = 1;
        A.best[a] = s;
        A.next_best[a] = A.nref > 1 ? s : neg_inf();   /* a second barcode ties: not > best, but > -Inf */
    }
}

/* ------------------------------------------------------------------------------------------------
 * Traceback: one thread per alignment walks the 4-bit records.
 *
 * Equivalence with reference backtrack<> (src/reference_align.cpp:231-278): a positive direction d at
 * (i,c) means "move left d columns"; d = 1 + pos - left_jump_points[i] counts the columns since the
 * running horizontal gap of row i was last re-opened, i.e. the run of consecutive E-extend decisions
 * (:134) ending at c.  p1 is that decision evaluated with the open penalty; the reference evaluates
 * it with the extend penalty when the left cell chose "left", in which case it is always false
 * (H == E bitwise there) -- hence ext = p1 && left-neighbour-did-not-choose-left.  Same for up/F.
 */
struct FlagReader {
    const TraceArgs& T;
    long long a;
    int len;
    const int* colinfo;   /* wavefront / solo layouts: per DP column, (word offset j*(kSkew*G+1)) << 8 | (bit shift 4k) */
    const void* flags;    /* the record planes of the source this alignment reads (TraceArgs::sel) */
    const void* flags_hi;
    __device__ unsigned get(int i, int c) const {
        if (T.layout != 1) {
            const int ci = colinfo[c];
            /* wavefront: word (i + kSkew*j)*G + j of the alignment; solo: [32 alignments][row][lane] */
            const long long w = (T.layout == 0) ? a * T.fstride + (long long)i * T.G + (ci >> 8)
                                                : ((a >> 5) * T.fstride + i) * 32 + (a & 31);
            const int sh = ci & 0xff;
            if (sh >= 64) {       /* columns 16.. of the lane live in the second plane */
                const unsigned v = (T.hi_bytes == 1) ? (unsigned)reinterpret_cast<const uint8_t*>(flags_hi)[w]
                                                     : reinterpret_cast<const uint32_t*>(flags_hi)[w];
                return (v >> (sh - 64)) & 15u;
            }
            /* the low word is wordbytes wide; slot k lives in its 32-bit part k/8 */
            const uint32_t* base = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(flags) + w * T.wordbytes);
            return (base[sh >> 5] >> (sh & 31)) & 15u;
        }
        return reinterpret_cast<const uint8_t*>(flags)[a * T.fstride + (long long)(c - 1) * len + (i - 1)];
    }
};

__global__ void __launch_bounds__(128) traceback(const TraceArgs T)
{
    extern __shared__ int tb_colinfo[];
    if (T.layout != 1) {
        const int pad = T.G * T.C - T.L;
        for (int c = 1 + threadIdx.x; c <= T.L; c += blockDim.x) {
            int j, k;
            wf_locate(c, T.C, pad, &j, &k);
            tb_colinfo[c] = ((j * (kSkew * T.G + 1)) << 8) | (4 * k);
        }
        __syncthreads();
    }
    const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= T.n) return;
    /* fused both-ends runs: the records of (adaptor, front window) and (adaptor, back window) are both on the device and
     * only the strand .resolve_strand kept is walked (sel[a] != 0: the second source) */
    const bool alt = T.sel != nullptr && T.sel[a] != 0;
    const int L = T.L;
    const int len = alt ? T.lens2[a] : T.lens[a];
    const int32_t* endrow = alt ? T.endrow2 : T.endrow;
    const int32_t* posv = alt ? T.pos2 : T.pos;
    const long long rec = posv ? (long long)posv[a] : a;       /* where this read's records are (speculative runs) */
    const long long n = T.n;
    const long long pitch = T.out_pitch > 0 ? T.out_pitch : n;
    FlagReader R{T, rec, len, tb_colinfo, alt ? T.flags2 : T.flags, alt ? T.flags_hi2 : T.flags_hi};
    int32_t* map = T.map + a;        /* map[c*n]: (row << 1) | is_match, fill_map's mapping (:280-305) */
    uint8_t* ops = T.ops ? T.ops + a * T.ops_stride : nullptr;
    int nops = 0;

    /* Flat three-state walk, one record read and one move per iteration (keeps the warp's lanes in step):
     *   ST_H  look at the cell's own choice;
     *   ST_E  inside a horizontal run: the move into this cell was "left", and whether the run continues was
     *         decided by the p1 bit of the cell we came from AND this cell not having chosen "left" itself
     *         (then its own choice says "left" again anyway) -- resolved when this cell's record is read;
     *   ST_F  same for vertical runs with p2 / "up". */
    enum { ST_H = 0, ST_E = 1, ST_F = 2 };
    int i = len, c = L, state = ST_H;
    bool pending = false;   /* p1 (in ST_E) / p2 (in ST_F) of the cell the current run came from */
    if (endrow && len > 0 && L > 0) {
        /* the forward kernel already followed the climb up the last column (land_step): skip those up-moves */
        i = endrow[rec];
        for (int x = i; x < len; ++x) {
            if (ops) ops[nops] = 'I';
            ++nops;
        }
    }
    while (c > 0) {
        unsigned f;
        int ch;
        if (i == 0) {          /* row 0: +1 (:118), never an extension */
            f = 0;
            ch = CH_LEFT;
        } else {
            f = R.get(i, c);
            ch = (f & 1u) ? CH_DIAG : ((f & 2u) ? CH_LEFT : CH_UP);
        }
        int move;
        if (i == 0) {
            move = CH_LEFT;    /* the reference's runs stop at row 0 / column 0 regardless of the extend bit */
        } else if (state == ST_E && pending && ch != CH_LEFT) {
            move = CH_LEFT;    /* the horizontal run passes through this cell whatever it chose */
        } else if (state == ST_F && pending && ch != CH_UP) {
            move = CH_UP;
        } else {
            move = ch;
        }
        if (move == CH_DIAG) {
            map[(long long)c * n] = (i << 1) | 1;
            if (ops) ops[nops] = 'M';
            ++nops;
            --i;
            --c;
            state = ST_H;
        } else if (move == CH_LEFT) {
            map[(long long)c * n] = ((i + 1) << 1);
            if (ops) ops[nops] = 'D';
            ++nops;
            pending = (f & 4u) != 0;
            --c;
            state = ST_E;
        } else {
            if (ops) ops[nops] = 'I';
            ++nops;
            pending = (f & 8u) != 0;
            --i;
            state = ST_F;
        }
    }
    while (i > 0) {   /* leading read bases, consumed at column 0 (:273-276) */
        if (ops) ops[nops] = 'I';
        ++nops;
        --i;
    }
    if (T.nops) T.nops[a] = nops;

    if (!T.start) return;
    /* querymap::operator() (:307-351) and the guards of src/adaptor_align.cpp:57-68 */
    const int nrows = len + 1;
    {
        const int m1 = map[n], mL = map[(long long)L * n];
        const int first = (m1 >> 1) - 1;
        const int second = (mL >> 1) + (mL & 1) - 1;
        int st = 0, en = 0;
        if (first < second) {
            st = first + 1;
            en = second;
        }
        if (T.width) {     /* adaptor2 into read coordinates: width - x + 1 (R/adaptorAlign.R:66-71; unaligned 0 -> width + 1) */
            const int w = T.width[a];
            st = w - st + 1;
            en = w - en + 1;
        }
        T.start[a] = st;
        T.end[a] = en;
    }
    for (int s = 0; s < T.nsec; ++s) {
        const int rs = T.sec_starts[s];
        int re = T.sec_ends[s];
        int curstart, curend;
        if (rs == 0) {
            curstart = 1;
        } else {
            const int m = map[(long long)rs * n];
            curstart = (m >> 1) + (m & 1);
        }
        ++re;
        if (re == L + 1) {
            curend = nrows;
        } else {
            curend = map[(long long)re * n] >> 1;
        }
        T.sec_start[(long long)s * pitch + a] = curstart;        /* (curstart-1) + 1 */
        T.sec_width[(long long)s * pitch + a] = curend - curstart;
    }
}

/* .resolve_strand (R/adaptorAlign.R:112-122) on four score vectors: fscore = pmax(s_a1_front, 0) + pmax(s_a2_back, 0),
 * rscore likewise on the swapped windows, reversed = fscore < rscore (strict), and
 * the two scores R keeps (ifelse(is.reverse, revcomp, forward), R/getAdaptorThresholds.R:123-127 / R/adaptorAlign.R:192-196).
 * Run between the forward passes and the tracebacks of a fused both-ends run, so that only the kept strand is walked. */
__global__ void resolve_strand(const StrandArgs S)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.n) return;
    const double s1 = S.a1_front[i], s2 = S.a2_back[i], r1 = S.a1_back[i], r2 = S.a2_front[i];
    const double f = __dadd_rn(fmax(s1, 0.0), fmax(s2, 0.0));
    const double r = __dadd_rn(fmax(r1, 0.0), fmax(r2, 0.0));
    const bool rev = f < r;
    if (S.reversed) S.reversed[i] = rev ? 1 : 0;
    if (S.score1) S.score1[i] = rev ? r1 : s1;
    if (S.score2) S.score2[i] = rev ? r2 : s2;
    if (S.strand_score) S.strand_score[i] = rev ? r : f;
    if (S.predicted) {
        const unsigned p = S.predicted[i];
        if (!rev && p == 1u) {            /* predicted reverse, kept forward: its forward-strand passes wrote no records */
            const int at = atomicAdd(S.ranges + 9, 1);
            S.list_fwd[at] = (int32_t)i;
            S.pos_fwd[i] = at;
        } else if (rev && p == 0u) {
            const int at = atomicAdd(S.ranges + 11, 1);
            S.list_rev[at] = (int32_t)i;
            S.pos_rev[i] = at;
        }
    }
}

/* See StrandLists (kernels.h).  Exact 8-mer seeds of adaptor1 / adaptor2 counted in the first `scan` bases of both
 * windows.  Four lanes share a read: lane k counts the 8-mers ENDING in bases [32k, 32k + 32) (+128, ...) of each window,
 * for which it also looks at the seven bases before them; eight reads per warp, warps stride over the launch. */
__device__ __forceinline__ void count_seeds(const uint16_t* row, int len, int s, const uint32_t* tab, int& h1, int& h2) {
    unsigned code = 0;
    int valid = 0;
    const int lo = s >= 7 ? s - 7 : 0, hi = min(len, s + 32);
    for (int i = lo; i < hi; ++i) {
        const int o = __ffs((unsigned)row[i] >> 8);             /* 1..4 for A, C, G, T; 0 otherwise */
        code = ((code << 2) | (unsigned)(o > 0 ? o - 1 : 0)) & 0xFFFFu;
        valid = o > 0 ? valid + 1 : 0;
        if (valid >= 8 && i >= s) {
            const unsigned t = tab[code >> 4] >> ((code & 15u) * 2u);
            h1 += t & 1u;
            h2 += (t >> 1) & 1u;
        }
    }
}

/* the same from 16-byte loads (rows 16-byte aligned, pitch a multiple of 8 entries) */
__device__ __forceinline__ void count_seeds_vec(const uint16_t* row, int len, int s, const uint32_t* tab, int& h1, int& h2) {
    uint32_t w[20];                                    /* bases s-8 .. s+31, two per word */
    const uint4* q = reinterpret_cast<const uint4*>(row + s);
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    const uint4 pre = s > 0 ? __ldg(q - 1) : zero;
    w[0] = pre.x; w[1] = pre.y; w[2] = pre.z; w[3] = pre.w;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint4 v = (s + 8 * k < len) ? __ldg(q + k) : zero;
        w[4 + 4 * k] = v.x; w[5 + 4 * k] = v.y; w[6 + 4 * k] = v.z; w[7 + 4 * k] = v.w;
    }
    unsigned code = 0;
    int valid = 0;
#pragma unroll
    for (int jj = 1; jj < 40; ++jj) {
        const int i = s - 8 + jj;
        const unsigned e = (jj & 1) ? (w[jj >> 1] >> 16) : (w[jj >> 1] & 0xFFFFu);
        const int o = (i >= 0 && i < len) ? __ffs(e >> 8) : 0;
        code = ((code << 2) | (unsigned)(o > 0 ? o - 1 : 0)) & 0xFFFFu;
        valid = o > 0 ? valid + 1 : 0;
        if (jj >= 8 && valid >= 8) {
            const unsigned t = tab[code >> 4] >> ((code & 15u) * 2u);
            h1 += t & 1u;
            h2 += (t >> 1) & 1u;
        }
    }
}

__global__ void __launch_bounds__(256) classify_strands(const ClassifyArgs A)
{
    __shared__ uint32_t tab[4096];          /* two bits per 8-mer: bit 0 = occurs in adaptor1, bit 1 = in adaptor2 */
    for (int x = threadIdx.x; x < 4096; x += blockDim.x) tab[x] = A.seeds[x];
    __syncthreads();
    const unsigned lane = threadIdx.x & 31, sub = lane & 3, slot = lane >> 2;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const bool vec = A.vec != 0;
    for (long long base = warp * 8; base < A.n; base += warps * 8) {
        const long long r = base + slot;
        const bool live = r < A.n;
        int fwd = 0, rev = 0;          /* adaptor1 x front + adaptor2 x back  vs  adaptor1 x back + adaptor2 x front */
        if (live) {
            const uint16_t* rf = A.rows_front + r * (long long)A.stride;
            const uint16_t* rb = A.rows_back + r * (long long)A.stride;
            const int lf = min(A.lens_front[r], A.scan), lb = min(A.lens_back[r], A.scan);     /* the adaptors sit at the start of their windows */
            for (int s = (int)sub * 32; s < lf; s += 128) {
                if (vec) count_seeds_vec(rf, lf, s, tab, fwd, rev);
                else count_seeds(rf, lf, s, tab, fwd, rev);
            }
            for (int s = (int)sub * 32; s < lb; s += 128) {
                if (vec) count_seeds_vec(rb, lb, s, tab, rev, fwd);
                else count_seeds(rb, lb, s, tab, rev, fwd);
            }
        }
#pragma unroll
        for (int d = 1; d < 4; d <<= 1) {
            fwd += __shfl_xor_sync(FULL, fwd, d);
            rev += __shfl_xor_sync(FULL, rev, d);
        }
        unsigned pred = fwd >= rev + A.margin ? 0u : (rev >= fwd + A.margin ? 1u : 2u);
        if (A.test_mode == 1 && pred != 2u) pred ^= 1u;
        if (A.test_mode == 2) pred = 2u;
        const bool mine = live && sub == 0;
        if (mine) A.L.predicted[r] = (uint8_t)pred;
        /* append to the four lists, one atomic per list per warp */
        auto append = [&](bool want, int32_t* list, int32_t* pos, int counter) {
            const unsigned m = __ballot_sync(FULL, want);
            if (m == 0) return;
            int at0 = 0;
            if (lane == (unsigned)(__ffs(m) - 1)) at0 = atomicAdd(A.L.ranges + counter, __popc(m));
            at0 = __shfl_sync(FULL, at0, __ffs(m) - 1);
            if (want) {
                const int at = at0 + __popc(m & ((1u << lane) - 1u));
                list[at] = (int32_t)r;
                if (pos) pos[r] = at;
            }
        };
        append(mine && pred != 1u, A.L.list_fwd, A.L.pos_fwd, 1);      /* forward-strand passes with records: predicted forward or unsure */
        append(mine && pred != 0u, A.L.list_rev, A.L.pos_rev, 3);
        append(mine && pred == 1u, A.L.list_sfwd, nullptr, 5);          /* forward-strand passes score-only */
        append(mine && pred == 0u, A.L.list_srev, nullptr, 7);
    }
}

/* The re-run ranges start where the predicted lists end. */
__global__ void finish_strand_lists(int32_t* ranges)
{
    ranges[8] = ranges[9] = ranges[1];
    ranges[10] = ranges[11] = ranges[3];
}

/* ---- counter-based randomness shared by the scramble and the synthetic-read generator ----------------------------
 * 32-bit integer hashing only (two multiplies per word), so that sarlacc_b200/synth.py and api.py reproduce every draw
 * bit for bit with numpy uint32 arithmetic.  A stream is keyed by (seed, global read index, field); word p of it is
 * hash32(key ^ (p * 0x9E3779B1 + 0x7F4A7C15)). */
__host__ __device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x21F0AAADu;
    x ^= x >> 15;
    x *= 0x735A2D97u;
    x ^= x >> 15;
    return x;
}

__host__ __device__ __forceinline__ uint32_t stream_key(unsigned long long seed, unsigned long long index, uint32_t field) {
    uint32_t k = hash32((uint32_t)seed ^ 0x243F6A88u);
    k = hash32(k ^ (uint32_t)(seed >> 32));
    k = hash32(k ^ (uint32_t)index);
    k = hash32(k ^ (uint32_t)(index >> 32));
    return hash32(k ^ (field * 0x85EBCA6Bu + 0xC2B2AE35u));
}

__host__ __device__ __forceinline__ uint32_t stream_word(uint32_t key, uint32_t p) {
    return hash32(key ^ (p * 0x9E3779B1u + 0x7F4A7C15u));
}

/* .scramble_input (R/getAdaptorThresholds.R:68-92) on packed rows: a Fisher-Yates shuffle of the window, bases and
 * qualities moving together (one 16-bit entry each).  for i = len-1 .. 1: j = floor(word_i * (i + 1) / 2^32); swap(i, j),
 * words from the stream keyed by (seed, read index, field 16 + stream_id): a uniform random permutation that depends on
 * nothing but those, so chunking, sharding and device count do not change it (api.py: _scramble_input is the host
 * mirror).  One thread per window, the window in shared memory as [entry][thread] (pitch TPB + 2 entries: the transposing
 * loads and stores of a warp then hit 32 different banks). */
template <int TPB>
__global__ void __launch_bounds__(TPB) scramble_rows_fy(const uint16_t* in, uint16_t* out, const int32_t* lens, long long n, int stride,
        unsigned long long seed, unsigned long long first_index, const unsigned long long* read_index, unsigned stream_id)
{
    extern __shared__ uint16_t swin[];          /* [stride][TPB + 2] */
    constexpr int P = TPB + 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long base = (long long)blockIdx.x * TPB; base < n; base += (long long)gridDim.x * TPB) {
        /* load: each warp brings in its own 32 windows, 32 consecutive entries of one window per instruction */
        for (int w = 0; w < 32; ++w) {
            const long long a = base + warp * 32 + w;
            if (a >= n) break;
            const int len = lens[a];
            const uint16_t* src = in + a * (long long)stride;
            for (int e = lane; e < len; e += 32) swin[e * P + warp * 32 + w] = src[e];
        }
        __syncwarp();
        {
            const long long a = base + threadIdx.x;
            if (a < n) {
                const int len = lens[a];
                const unsigned long long rid = read_index ? read_index[a] : first_index + (unsigned long long)a;
                const uint32_t key = stream_key(seed, rid, 16u + stream_id);
                uint16_t* mine = swin + threadIdx.x;
                for (int i = len - 1; i >= 1; --i) {
                    const uint32_t j = __umulhi(stream_word(key, (uint32_t)i), (uint32_t)(i + 1));
                    const uint16_t x = mine[i * P], y = mine[j * P];
                    mine[i * P] = y;
                    mine[j * P] = x;
                }
            }
        }
        __syncwarp();
        for (int w = 0; w < 32; ++w) {
            const long long a = base + warp * 32 + w;
            if (a >= n) break;
            const int len = lens[a];
            uint16_t* dst = out + a * (long long)stride;
            for (int e = lane; e < stride; e += 32) dst[e] = e < len ? swin[e * P + warp * 32 + w] : (uint16_t)0;
        }
        __syncwarp();
    }
}

/* The same shuffle in place in global memory, for windows too long for the shared-memory kernel. */
__global__ void __launch_bounds__(128) scramble_rows_fy_global(const uint16_t* in, uint16_t* out, const int32_t* lens, long long n, int stride,
        unsigned long long seed, unsigned long long first_index, const unsigned long long* read_index, unsigned stream_id)
{
    const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const int len = lens[a];
    const uint16_t* src = in + a * (long long)stride;
    uint16_t* dst = out + a * (long long)stride;
    for (int e = 0; e < stride; ++e) dst[e] = e < len ? src[e] : (uint16_t)0;
    const unsigned long long rid = read_index ? read_index[a] : first_index + (unsigned long long)a;
    const uint32_t key = stream_key(seed, rid, 16u + stream_id);
    for (int i = len - 1; i >= 1; --i) {
        const uint32_t j = __umulhi(stream_word(key, (uint32_t)i), (uint32_t)(i + 1));
        const uint16_t x = dst[i], y = dst[j];
        dst[i] = y;
        dst[j] = x;
    }
}

/* ---- synthetic reads on the device: the mockReads generator (R/mockReads.R:58-92) ------------------------------------
 * One thread per (read, end).  The front window of an unflipped read is the first `tol` bases of the mutated molecule
 * adaptor1 (first N-run <- one random base repeated, :50, or a barcode; other ambiguous positions <- random bases) +
 * random insert; its back window (already reverse-complemented, as .get_front_and_back hands it over) is the first `tol`
 * bases of mutated adaptor2 + random insert (:58-64).  Per molecule base: substituted by a uniform base w.p. sub_rate
 * (:73-74), then w.p. indel_rate replaced by 0 or 2..max_insert copies of itself (:77-79).  Qualities are iid Phred of
 * U(0, max_err) through an integer threshold table built on the host (:82); half of the reads are flipped (:91-92), which
 * swaps the two windows.  Every draw is stream_word(stream_key(seed, read index, field), position): the read depends on
 * (seed, index) only -- sarlacc_b200/synth.py: mock_windows is the numpy mirror, bit for bit.
 * Fields: 0/4 molecule fill (front/back molecule), 1/5 mutation, 2/6 count choice, 3/7 quality, 8 flip, 9 width. */
/* quality >= k  <=>  word < thr[k] with thr non-increasing in k: the quality is the largest k in 0..93 whose threshold
 * exceeds the word, found by a fixed-depth binary search (every lane takes the same seven steps). */
__device__ __forceinline__ unsigned mock_quality_search(uint32_t h, const uint32_t* thr, unsigned qmin) {
    unsigned lo = 0, hi = 94;            /* invariant: h < thr[lo] (thr[0] = 2^32 - 1 stands for 2^32), !(h < thr[hi]) (thr[94] = 0) */
#pragma unroll
    for (int s = 0; s < 7; ++s) {
        const unsigned mid = (lo + hi) >> 1;
        const bool below = mid > lo && h < thr[mid];
        lo = below ? mid : lo;
        hi = below ? hi : (mid > lo ? mid : hi);
    }
    return lo < qmin ? qmin : lo;        /* word 2^32 - 1 fails the saturated thresholds below qmin */
}

/* The same function through a 256-entry table on the word's top byte: lut[b] = the quality of the largest word of bucket
 * b if the bucket spans at most two qualities (then one comparison decides), 255 otherwise (the search; only the lowest
 * buckets, where the thresholds crowd: < 1 % of the words). */
__device__ __forceinline__ unsigned mock_quality(uint32_t h, const uint32_t* thr, const uint8_t* lut, unsigned qmin) {
    const unsigned base = lut[h >> 24];
    if (base == 255u) return mock_quality_search(h, thr, qmin);
    return base + (h < thr[base + 1] ? 1u : 0u);
}

/* Threads [0, n) build the adaptor1 molecule's window of read t, threads [n, 2n) the adaptor2 molecule's: a warp works on
 * one molecule (uniform adaptor reads); the quality thresholds and their lookup table sit in shared memory. */
__global__ void __launch_bounds__(128) mock_windows_kernel(const __grid_constant__ MockArgs M)
{
    __shared__ uint32_t sthr[96];
    __shared__ uint32_t swcdf[256];
    __shared__ uint8_t slut[256];
    __shared__ char sad[2][128];
    for (int x = threadIdx.x; x < 95; x += blockDim.x) sthr[x] = M.qthr[x];
    for (int x = threadIdx.x; x < 256; x += blockDim.x) swcdf[x] = M.wcdf[x];
    for (int x = threadIdx.x; x < 128; x += blockDim.x) { sad[0][x] = M.adaptor1[x]; sad[1][x] = M.adaptor2[x]; }
    __syncthreads();
    for (int x = threadIdx.x; x < 256; x += blockDim.x) {
        const unsigned qlo = mock_quality_search(((uint32_t)x << 24) | 0xFFFFFFu, sthr, M.qmin);
        const unsigned qhi = mock_quality_search((uint32_t)x << 24, sthr, M.qmin);
        slut[x] = (qhi - qlo <= 1u && qlo < 93u) ? (uint8_t)qlo : (uint8_t)255;
    }
    __syncthreads();
    const long long npad = (M.n + 127) / 128 * 128;            /* molecule 1 starts at a block boundary */
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int mol = t >= npad ? 1 : 0;                          /* 0: adaptor1 molecule, 1: adaptor2 molecule */
    const long long r = mol ? t - npad : t;
    if (r >= M.n) return;
    const unsigned long long rid = M.first_index + (unsigned long long)r;
    const bool flip = (stream_word(stream_key(M.seed, rid, 8u), 0u) >> 31) != 0u;
    const bool to_front = (mol == 0) != flip;
    uint16_t* row = (to_front ? M.front : M.back) + r * (long long)M.stride;
    const char* ad = sad[mol];
    const int alen = mol == 0 ? M.len1 : M.len2;
    const int run0 = mol == 0 ? M.run0_start : -1, run1 = mol == 0 ? M.run0_end : -1;
    const uint32_t kfill = stream_key(M.seed, rid, 0u + 4u * mol), kmut = stream_key(M.seed, rid, 1u + 4u * mol);
    const uint32_t kcnt = stream_key(M.seed, rid, 2u + 4u * mol), kq = stream_key(M.seed, rid, 3u + 4u * mol);
    const unsigned barcode_base = stream_word(kfill, 0xFFFFFFFFu) & 3u;
    const unsigned barcode_pick = M.nbarcodes > 0 ? stream_word(kfill, 0xFFFFFFFEu) % (unsigned)M.nbarcodes : 0u;
    int emitted = 0;
    uint32_t acc[4] = {0u, 0u, 0u, 0u};                     /* eight 16-bit entries per 128-bit store */
    for (uint32_t p = 0; emitted < M.tol; ++p) {
        unsigned b;
        const char c = p < (uint32_t)alen ? ad[p] : 'N';        /* uniform over the warp: every lane is at molecule position p */
        if (c == 'A') b = 0; else if (c == 'C') b = 1; else if (c == 'G') b = 2; else if (c == 'T') b = 3;
        else if ((int)p >= run0 && (int)p < run1) b = M.nbarcodes > 0 ? (unsigned)M.barcodes[barcode_pick * (run1 - run0) + ((int)p - run0)] : barcode_base;
        else b = stream_word(kfill, p) & 3u;
        const uint32_t u = stream_word(kmut, p);
        if ((u >> 16) < M.sub_thr) b = (u >> 2) & 3u;       /* substitution: a uniform base (may be the same one) */
        int copies = 1;
        if ((u & 0xFFFFu) < M.indel_thr) {                  /* indel: 0 or 2..max_insert copies */
            const uint32_t k = stream_word(kcnt, p) % (uint32_t)M.max_insert;
            copies = k == 0 ? 0 : (int)k + 1;
        }
        for (int x = 0; x < copies && emitted < M.tol; ++x) {
            const unsigned q = mock_quality(stream_word(kq, (uint32_t)emitted), sthr, slut, M.qmin);
            const uint32_t entry = ((1u << b) << 8) | q;
            acc[(emitted & 7) >> 1] |= entry << (16 * (emitted & 1));
            ++emitted;
            if ((emitted & 7) == 0) {
                *reinterpret_cast<uint4*>(row + emitted - 8) = make_uint4(acc[0], acc[1], acc[2], acc[3]);
                acc[0] = acc[1] = acc[2] = acc[3] = 0u;
            }
        }
    }
    /* tail of the last 8-entry group and zero padding up to the row stride */
    for (int e = emitted & ~7; e < M.stride; e += 8) {
        *reinterpret_cast<uint4*>(row + e) = make_uint4(acc[0], acc[1], acc[2], acc[3]);
        acc[0] = acc[1] = acc[2] = acc[3] = 0u;
    }
    if (mol == 0) {
        M.lens_front[r] = M.tol;
        M.lens_back[r] = M.tol;
        if (M.flipped) M.flipped[r] = flip ? 1 : 0;
        /* read width = mutated length of adaptor1 + insert + adaptor2 (:77-79 over the whole molecule): the number of
         * indels is Binomial(molecule_len, indel rate), drawn by inverting its distribution function (an integer table
         * built on the host: E = wlo + #{k : word >= wcdf[k]}); each changes the length by -1 or +1..max_insert-1 */
        const uint32_t kw = stream_key(M.seed, rid, 9u);
        const uint32_t u = stream_word(kw, 0u);
        unsigned lo = 0, hi = 256;                             /* first k with u < wcdf[k] (256 if none) */
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const unsigned mid = (lo + hi) >> 1;
            const bool ge = u >= swcdf[mid];
            lo = ge ? mid + 1 : lo;
            hi = ge ? hi : mid;
        }
        const int events = M.wlo + (int)lo;
        int width = M.molecule_len;
        for (int e = 0; e < events; ++e) {
            const uint32_t k = stream_word(kw, 1u + (uint32_t)e) % (uint32_t)M.max_insert;
            width += k == 0 ? -1 : (int)k;
        }
        M.width[r] = width;
    }
}

/* One warp per window, lanes across its bases: coalesced byte loads, 64-byte row stores; tables in shared memory. */
__global__ void __launch_bounds__(256) pack_rows_kernel(const __grid_constant__ PackArgs A)
{
    __shared__ uint8_t sbase[256];
    __shared__ uint16_t sq[256];
    const uint8_t* bt = A.back ? A.base_rc : A.base;
    for (int x = threadIdx.x; x < 256; x += blockDim.x) {
        sbase[x] = bt[x];
        sq[x] = A.qidx[x];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < A.n; i += warps) {
        const int len = A.lens[i];
        const long long so = A.soff[i];
        const uint8_t* s = A.seq + so;
        const uint8_t* q = A.qual + A.qoff[i];
        uint16_t* out = A.rows + i * (long long)A.stride;
        bool bad = false;
        for (int r = lane; r < len; r += 32) {
            const int src = A.back ? len - 1 - r : r;
            unsigned qi = sq[q[src]];
            if (qi & 0xFF00u) { bad = true; qi = 0; }
            unsigned code;
            if (A.seq4) {       /* the host sent the forward table's codes; the complement of a one-hot code is its bit reversal */
                const long long at = so + src;
                code = ((unsigned)A.seq[at >> 1] >> (((unsigned)at & 1u) * 4u)) & 15u;
                if (A.back) code = __brev(code) >> 28;
            } else {
                code = sbase[s[src]];
            }
            out[r] = (uint16_t)(qi | (code << 8));
        }
        if (A.first_bad && __any_sync(FULL, bad) && lane == 0) atomicMin(A.first_bad, i);
    }
}

template <int C, bool TRACE, bool SOLO>
auto wf_kernel(bool pair) -> void (*)(const AlignArgs) {
    if constexpr (SOLO) {
        return wf_forward2<C, TRACE, true>;      /* the solo geometry exists as a row-pair kernel only */
    } else {
        return pair ? wf_forward2<C, TRACE, false> : wf_forward<C, TRACE>;
    }
}

template <int C, bool TRACE, bool SOLO>
int wf_resident_blocks(bool pair, size_t smem) {
    auto kern = wf_kernel<C, TRACE, SOLO>(pair);
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kern, kBlock, smem);
    return per < 1 ? 1 : per;
}

template <int C, bool TRACE, bool SOLO>
const char* launch_wf(const AlignArgs& a, int grid, cudaStream_t st, size_t smem) {
    const bool pair = SOLO || a.pair_rows != 0;
    auto kern = wf_kernel<C, TRACE, SOLO>(pair);
    if (grid <= 0) {
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        grid = sms * wf_resident_blocks<C, TRACE, SOLO>(pair, smem);
    } else if (smem > 48 * 1024) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    /* never more groups than alignments */
    const long long groups_per_block = (long long)(kBlock / 32) * (32 / a.G);
    const long long need = (a.n + groups_per_block - 1) / groups_per_block;
    if (need < grid) grid = (int)(need > 0 ? need : 1);
    kern<<<grid, kBlock, smem, st>>>(a);
    if (SOLO) return TRACE ? "solo_forward2<trace>" : "solo_forward2<score>";
    return pair ? (TRACE ? "wf_forward2<trace>" : "wf_forward2<score>") : (TRACE ? "wf_forward<trace>" : "wf_forward<score>");
}

/* Calls f(std::integral_constant<int, C>, std::integral_constant<bool, SOLO>) for an instantiated geometry; false if
 * (C, solo) is not one. */
template <class F>
bool for_geometry(int C, bool solo, F&& f) {
#ifdef SARLACC_ONLY_C   /* tuning builds: instantiate one geometry only (tools/variants.py) */
    if (C == SARLACC_ONLY_C) {
        if (solo) {
            if constexpr (SARLACC_ONLY_C >= kSoloMinC && SARLACC_ONLY_C <= kSoloMaxC) { f(std::integral_constant<int, SARLACC_ONLY_C>(), std::true_type()); return true; }
        } else {
            if constexpr (SARLACC_ONLY_C <= kMaxC) { f(std::integral_constant<int, SARLACC_ONLY_C>(), std::false_type()); return true; }
        }
    }
    return false;
#else
    if (solo) {
        switch (C) {
#define SARLACC_GEOM(N) case N: f(std::integral_constant<int, N>(), std::true_type()); return true;
            SARLACC_GEOM(20) SARLACC_GEOM(21) SARLACC_GEOM(22) SARLACC_GEOM(23) SARLACC_GEOM(24)
#undef SARLACC_GEOM
        }
        return false;
    }
    switch (C) {
#define SARLACC_GEOM(N) case N: f(std::integral_constant<int, N>(), std::false_type()); return true;
        SARLACC_GEOM(1) SARLACC_GEOM(2) SARLACC_GEOM(3) SARLACC_GEOM(4) SARLACC_GEOM(5) SARLACC_GEOM(6)
        SARLACC_GEOM(7) SARLACC_GEOM(8) SARLACC_GEOM(9) SARLACC_GEOM(10) SARLACC_GEOM(11) SARLACC_GEOM(12)
        SARLACC_GEOM(14) SARLACC_GEOM(16) SARLACC_GEOM(18)
#undef SARLACC_GEOM
    }
    return false;
#endif
}

}  // namespace

size_t wavefront_smem_bytes(const AlignArgs& a) {
    /* row-0 chain + lane-private tables (two rows' worth: the row-pair kernel) + per-quality cost tables (5 doubles) + reference columns */
    return sizeof(double) * ((((size_t)a.L + 1 + (size_t)(kBlock / 32) * 2 * kCostEntries * 32 + 1) & ~(size_t)1) +
                             (size_t)a.enc_n * 5 + 2) +
           2 * (size_t)a.nref * a.L;
}

int wavefront_block_threads() { return kBlock; }

const char* launch_wavefront(const AlignArgs& a, bool trace, bool has_alt, int grid, cudaStream_t st) {
    (void)has_alt;
    const size_t smem = wavefront_smem_bytes(a);
    const char* name = nullptr;
    for_geometry(a.C, a.solo != 0, [&](auto ctag, auto stag) {
        constexpr int C = decltype(ctag)::value;
        constexpr bool SOLO = decltype(stag)::value;
        name = trace ? launch_wf<C, true, SOLO>(a, grid, st, smem) : launch_wf<C, false, SOLO>(a, grid, st, smem);
    });
    return name;
}

long long wavefront_groups(const AlignArgs& a, bool trace) {
    const size_t smem = wavefront_smem_bytes(a);
    int per = 0;
    for_geometry(a.C, a.solo != 0, [&](auto ctag, auto stag) {
        constexpr int C = decltype(ctag)::value;
        constexpr bool SOLO = decltype(stag)::value;
        const bool pair = SOLO || a.pair_rows != 0;
        per = trace ? wf_resident_blocks<C, true, SOLO>(pair, smem) : wf_resident_blocks<C, false, SOLO>(pair, smem);
    });
    if (per == 0) return 0;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return (long long)sms * per * (kBlock / 32) * (32 / a.G);
}

const char* launch_generic(const AlignArgs& a, bool trace, int grid, cudaStream_t st) {
    const int block = 128;
    if (grid <= 0) grid = (int)((a.gthreads + block - 1) / block);
    generic_forward<<<grid, block, 0, st>>>(a, trace ? 1 : 0);
    return trace ? "generic_forward<trace>" : "generic_forward<>";
}

void launch_traceback(const TraceArgs& t, cudaStream_t st) {
    const int block = 128;
    const int grid = (int)((t.n + block - 1) / block);
    if (grid > 0) traceback<<<grid, block, sizeof(int) * ((size_t)t.L + 1), st>>>(t);
}

void launch_scramble(const uint16_t* in, uint16_t* out, const int32_t* lens, long long n, int stride,
                     unsigned long long seed, unsigned long long first_index, const unsigned long long* read_index,
                     unsigned long long stream_id, cudaStream_t st)
{
    if (n <= 0) return;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    auto launch = [&](auto tag) {
        constexpr int TPB = decltype(tag)::value;
        const size_t smem = sizeof(uint16_t) * (size_t)stride * (TPB + 2);
        cudaFuncSetAttribute(scramble_rows_fy<TPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int per = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, scramble_rows_fy<TPB>, TPB, smem);
        long long grid = (n + TPB - 1) / TPB;
        if (grid > (long long)sms * (per < 1 ? 1 : per)) grid = (long long)sms * (per < 1 ? 1 : per);
        scramble_rows_fy<TPB><<<(int)grid, TPB, smem, st>>>(in, out, lens, n, stride, seed, first_index, read_index, (unsigned)stream_id);
    };
    const size_t budget = 200u << 10;
    if (sizeof(uint16_t) * (size_t)stride * (128 + 2) <= budget / 2) launch(std::integral_constant<int, 128>());      /* two blocks per SM */
    else if (sizeof(uint16_t) * (size_t)stride * (64 + 2) <= budget) launch(std::integral_constant<int, 64>());
    else if (sizeof(uint16_t) * (size_t)stride * (32 + 2) <= budget) launch(std::integral_constant<int, 32>());
    else scramble_rows_fy_global<<<(int)((n + 127) / 128), 128, 0, st>>>(in, out, lens, n, stride, seed, first_index, read_index, (unsigned)stream_id);
}

void launch_classify_strands(const ClassifyArgs& c, cudaStream_t st) {
    if (c.n <= 0) return;
    ClassifyArgs a = c;
    a.vec = (c.stride % 8 == 0 && (reinterpret_cast<uintptr_t>(c.rows_front) & 15) == 0 && (reinterpret_cast<uintptr_t>(c.rows_back) & 15) == 0) ? 1 : 0;
    const long long blocks = (c.n + 63) / 64;           /* 8 warps x 8 reads */
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long full = (long long)sms * 8;
    classify_strands<<<(int)(blocks < full ? blocks : full), 256, 0, st>>>(a);
    finish_strand_lists<<<1, 1, 0, st>>>(c.L.ranges);
}

void launch_resolve_strand(const StrandArgs& s, cudaStream_t st) {
    const int block = 256;
    const int grid = (int)((s.n + block - 1) / block);
    if (grid > 0) resolve_strand<<<grid, block, 0, st>>>(s);
}

void launch_mock_windows(const MockArgs& m, cudaStream_t st) {
    if (m.n <= 0) return;
    const int block = 128;
    const long long npad = (m.n + block - 1) / block * block;
    mock_windows_kernel<<<(int)(2 * npad / block), block, 0, st>>>(m);
}

void launch_pack_rows(const PackArgs& a, cudaStream_t st) {
    if (a.n <= 0) return;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long grid = (a.n + 7) / 8;
    if (grid > (long long)sms * 16) grid = (long long)sms * 16;
    pack_rows_kernel<<<(int)grid, 256, 0, st>>>(a);
}

void launch_fill_empty(const AlignArgs& a, cudaStream_t st) {
    const int block = 256;
    const int grid = (int)((a.n + block - 1) / block);
    if (grid > 0) fill_empty<<<grid, block, 0, st>>>(a);
}

}  // namespace sarlacc
