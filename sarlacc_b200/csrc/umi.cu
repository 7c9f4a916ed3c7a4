/* UMI grouping on the device: replaces umi_group (src/umi_group.cpp:14-117) -- the trie-based
 * bounded Levenshtein search of src/sorted_trie.cpp and the greedy clustering of src/cluster_umis.cpp.
 *
 * What the reference computes, per pre-group of reads:
 *   1. for every UMI s, the stored UMIs t with lev2(s, t) <= 2 * threshold, where lev2 is the edit distance with
 *      indel = mismatch = 2 and "N against anything" = 1 (get_edit_score, src/sorted_trie.cpp:15-21), listed in the
 *      order the trie walk meets them: sorted by (sequence under A<C<G<T<N, shorter prefix first, insertion index)
 *      (src/sorted_trie.cpp:152-185); with two UMIs, the UMI2 list filtered by membership in the UMI1 list
 *      (src/umi_group.cpp:65-101);
 *   2. greedy clustering of those lists (cluster_umis, src/cluster_umis.cpp:7-112).
 *
 * Here step 1 is an all-pairs pass on the GPU: the reads of every pre-group are sorted once into trie order on the
 * host, one thread takes one query UMI and walks its group's candidates in that order, running the reference's DP
 * row by row in registers (row = one candidate base, columns = query positions) with the trie's own pruning rule
 * as early exit (a row whose minimum exceeds the limit cannot come back, src/sorted_trie.cpp:160-176).  Two passes:
 * count, exclusive scan on the host, fill -- so the lists come out dense, in order, without atomics.
 * Step 2 is inherently sequential (each pick changes the counts the next pick depends on) and stays on the host,
 * with a lazy max-heap instead of the reference's O(n) scan per cluster; ties resolve identically (largest index).
 */
#include "sarlacc_b200.h"
#include "kernels.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <queue>
#include <string>
#include <vector>

namespace sarlacc {
int set_error(const std::string& msg);
void count_launches(int n);
}

namespace {

constexpr int kUmiBlock = 128;

struct UmiArgs {
    const uint8_t* seq1;      /* [E][W1] candidate/query bytes, zero padded */
    const uint8_t* seq2;      /* [E][W2] or null */
    const uint8_t* len1;      /* [E] */
    const uint8_t* len2;
    const uint8_t* flags;     /* bit 0: UMI1 is in the trie (ACGTN only), bit 1: UMI2 is */
    const int32_t* gstart;    /* [E] candidate range of the entry's pre-group, in sorted positions */
    const int32_t* gend;
    const int32_t* local;     /* [E] index within the pre-group (what the reference stores in its lists) */
    int W1, W2;
    int limit1, limit2;       /* already doubled: 2 * threshold */
    long long E;
    int32_t* count;           /* pass 1 out */
    int32_t* first;           /* pass 1 out: [E][kUmiCap] the first matches of every read */
    const long long* offset;  /* pass 2 in */
    int32_t* neighbors;       /* pass 2 out */
};

/* lev2(query, cand) <= limit ?  Query bytes sit in registers (MAXQ / 4 words), the DP row too. */
template <int MAXQ>
__device__ __forceinline__ bool within(const uint32_t* qw, int lq, const uint8_t* cand, int lt, int limit) {
    const int dl = lq > lt ? lq - lt : lt - lq;
    if (2 * dl > limit) return false;
    int row[MAXQ + 1];
#pragma unroll
    for (int i = 0; i <= MAXQ; ++i) row[i] = 2 * i;                       /* src/sorted_trie.cpp:196-202 */
    for (int d = 0; d < lt; ++d) {
        const unsigned tb = cand[d];
        const bool tn = tb == 'N';
        int diag = row[0];
        row[0] = diag + 2;                                                 /* :111 */
        int rmin = row[0];
#pragma unroll
        for (int i = 1; i <= MAXQ; ++i) {
            const unsigned qb = (qw[(i - 1) >> 2] >> (8 * ((i - 1) & 3))) & 0xffu;
            const int sc = (tn || qb == 'N') ? 1 : (qb == tb ? 0 : 2);    /* get_edit_score, :15-21 */
            const int up = row[i];
            const int v = min(min(up + 2, row[i - 1] + 2), diag + sc);    /* :134-138 */
            diag = up;
            row[i] = v;
            rmin = min(rmin, v);
        }
        /* cells past the query's end only ever derive from real ones plus costs, so a row minimum above the limit
         * is final (the trie's pruning rule, :160-176) */
        if (rmin > limit) return false;
    }
    int res = 0;
#pragma unroll
    for (int i = 0; i <= MAXQ; ++i) res = (i == lq) ? row[i] : res;
    return res <= limit;
}

/* The same decision inside the Ukkonen band |i - d| <= K, K = threshold: a cell further from the diagonal has paid more
 * than K indels of cost 2 and is already over the limit, so leaving it at "infinity" changes no answer.  The band of
 * the previous row sits in 2K+1 registers; the query, shifted by one base per candidate base, in MAXQ/4 + 2 words, so
 * that every index is a compile-time constant. */
template <int MAXQ, int K>
__device__ __forceinline__ bool within_band(const uint32_t* qw, int lq, const uint8_t* cand, int lt, int limit) {
    const int dl = lq > lt ? lq - lt : lt - lq;
    if (2 * dl > limit) return false;
    constexpr int INF = 1 << 20;
    constexpr int NQ = MAXQ / 4;
    constexpr int NW = NQ + 2;
    /* s = K zero bytes, the query, zeros: byte m of s is the query base of band slot m in the row being computed */
    uint32_t sreg[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        const uint32_t lo = (w - K / 4 - 1 >= 0 && w - K / 4 - 1 < NQ) ? qw[w - K / 4 - 1] : 0u;
        const uint32_t hi = (w - K / 4 >= 0 && w - K / 4 < NQ) ? qw[w - K / 4] : 0u;
        sreg[w] = (K % 4 == 0) ? hi : __funnelshift_l(lo, hi, 8 * (K % 4));
    }
    int band[2 * K + 1];
#pragma unroll
    for (int m = 0; m <= 2 * K; ++m) band[m] = (m >= K) ? 2 * (m - K) : INF;     /* row 0: cell (0, i) = 2i */
    for (int d = 0; d < lt; ++d) {
        const unsigned tb = cand[d];
        const bool tn = tb == 'N';
        int left = INF, rmin = INF;
#pragma unroll
        for (int m = 0; m <= 2 * K; ++m) {
            const unsigned qb = (sreg[m >> 2] >> (8 * (m & 3))) & 0xffu;
            const int sc = (tn || qb == 'N') ? 1 : (qb == tb ? 0 : 2);
            const int diag = band[m];
            const int up = (m < 2 * K) ? band[m + 1] : INF;
            const int v = min(min(up + 2, left + 2), diag + sc);
            band[m] = v;
            left = v;
            rmin = min(rmin, v);
        }
        if (rmin > limit) return false;
#pragma unroll
        for (int w = 0; w < NW - 1; ++w) sreg[w] = __funnelshift_r(sreg[w], sreg[w + 1], 8);
        sreg[NW - 1] >>= 8;
    }
    int res = INF;
#pragma unroll
    for (int m = 0; m <= 2 * K; ++m) res = (m - K == lq - lt) ? band[m] : res;
    return res <= limit;
}

template <int MAXQ>
__device__ __forceinline__ bool within_any(const uint32_t* qw, int lq, const uint8_t* cand, int lt, int limit) {
    switch (limit >> 1) {     /* limit = 2 * threshold */
        case 0: return within_band<MAXQ, 0>(qw, lq, cand, lt, limit);
        case 1: return within_band<MAXQ, 1>(qw, lq, cand, lt, limit);
        case 2: return within_band<MAXQ, 2>(qw, lq, cand, lt, limit);
        case 3: return within_band<MAXQ, 3>(qw, lq, cand, lt, limit);
        case 4: return within_band<MAXQ, 4>(qw, lq, cand, lt, limit);
        default: return within<MAXQ>(qw, lq, cand, lt, limit);
    }
}

constexpr int kUmiCap = 16;   /* matches kept by the counting pass; only reads with more are walked a second time */

template <int MAXQ1, int MAXQ2, bool FILL>
__global__ void __launch_bounds__(kUmiBlock) umi_neighbors(const UmiArgs A) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= A.E) return;
    uint32_t q1[MAXQ1 / 4], q2[MAXQ2 / 4];
#pragma unroll
    for (int w = 0; w < MAXQ1 / 4; ++w) q1[w] = (4 * w < A.W1) ? reinterpret_cast<const uint32_t*>(A.seq1 + e * A.W1)[w] : 0u;
    const int l1 = A.len1[e];
    int l2 = 0;
    if (A.seq2) {
#pragma unroll
        for (int w = 0; w < MAXQ2 / 4; ++w) q2[w] = (4 * w < A.W2) ? reinterpret_cast<const uint32_t*>(A.seq2 + e * A.W2)[w] : 0u;
        l2 = A.len2[e];
    }
    const int need = A.seq2 ? 3 : 1;
    int n = 0;
    int32_t* out = FILL ? A.neighbors + A.offset[e] : A.first + e * kUmiCap;
    if (FILL) {
        const int have = A.count[e];
        if (have <= kUmiCap) {       /* the counting pass already holds this read's whole list */
            for (int k = 0; k < have; ++k) out[k] = A.first[e * kUmiCap + k];
            return;
        }
    }
    const int gs = A.gstart[e], ge = A.gend[e];
    for (int j = gs; j < ge; ++j) {
        if ((A.flags[j] & need) != need) continue;          /* never inserted into the trie (src/sorted_trie.cpp:53-69) */
        /* UMI2 first when present (its list drives the order, UMI1 filters) */
        if (A.seq2 && !within_any<MAXQ2>(q2, l2, A.seq2 + (long long)j * A.W2, A.len2[j], A.limit2)) continue;
        if (!within_any<MAXQ1>(q1, l1, A.seq1 + (long long)j * A.W1, A.len1[j], A.limit1)) continue;
        if (FILL || n < kUmiCap) out[n] = A.local[j];
        ++n;
    }
    if (!FILL) A.count[e] = n;
}

template <int M1, int M2>
void launch_pair(const UmiArgs& a, bool fill, cudaStream_t st) {
    const int grid = (int)((a.E + kUmiBlock - 1) / kUmiBlock);
    if (grid <= 0) return;
    if (fill) umi_neighbors<M1, M2, true><<<grid, kUmiBlock, 0, st>>>(a);
    else umi_neighbors<M1, M2, false><<<grid, kUmiBlock, 0, st>>>(a);
}

template <int M1>
bool launch_m2(const UmiArgs& a, bool fill, cudaStream_t st) {
    if (!a.seq2) { launch_pair<M1, 4>(a, fill, st); return true; }
    if (a.W2 <= 16) { launch_pair<M1, 16>(a, fill, st); return true; }
    if (a.W2 <= 32) { launch_pair<M1, 32>(a, fill, st); return true; }
    if (a.W2 <= 64) { launch_pair<M1, 64>(a, fill, st); return true; }
    return false;
}

bool launch_umi(const UmiArgs& a, bool fill, cudaStream_t st) {
    if (a.W1 <= 16) return launch_m2<16>(a, fill, st);
    if (a.W1 <= 32) return launch_m2<32>(a, fill, st);
    if (a.W1 <= 64) return launch_m2<64>(a, fill, st);
    return false;
}

struct Lists {   /* a list of integer vectors, flattened */
    std::vector<int64_t> off{0};
    std::vector<int32_t> values;
    void push(const std::vector<int32_t>& v) {
        values.insert(values.end(), v.begin(), v.end());
        off.push_back((int64_t)values.size());
    }
};

inline int trie_rank(uint8_t c) {   /* children order of the trie: src/sorted_trie.cpp:10,53-69 */
    switch (c) {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'T': return 3;
        case 'N': return 4;
        default: return 5;
    }
}

struct Seqs {
    const uint8_t* pool;
    const int64_t* off;
    const uint8_t* ptr(int64_t i) const { return pool + off[i]; }
    int64_t len(int64_t i) const { return off[i + 1] - off[i]; }
    bool storable(int64_t i) const {
        for (int64_t k = off[i]; k < off[i + 1]; ++k) if (trie_rank(pool[k]) == 5) return false;
        return true;
    }
    /* order of sorted_trie's walk: (sequence under ACGTN, prefix first) */
    bool less(int64_t a, int64_t b) const {
        const int64_t la = len(a), lb = len(b), m = std::min(la, lb);
        const uint8_t* pa = ptr(a);
        const uint8_t* pb = ptr(b);
        for (int64_t k = 0; k < m; ++k) {
            const int ra = trie_rank(pa[k]), rb = trie_rank(pb[k]);
            if (ra != rb) return ra < rb;
        }
        return la < lb;
    }
    bool same(int64_t a, int64_t b) const { return len(a) == len(b) && std::memcmp(ptr(a), ptr(b), (size_t)len(a)) == 0; }
};

/* cluster_umis (src/cluster_umis.cpp:7-112) over CSR lists of group-local indices.  Returns false with `msg` set on
 * the reference's two error conditions. */
bool cluster_lists(const long long* off, const int32_t* nb, int n, std::vector<std::vector<int32_t> >& out, const char** msg) {
    std::vector<int64_t> remaining((size_t)n);
    std::vector<char> in_play((size_t)n, 0);
    typedef std::pair<int64_t, int32_t> Key;   /* (remaining, index): the reference's max_element order, :62-69 */
    std::priority_queue<Key> heap;
    for (int a = 0; a < n; ++a) {
        const int64_t cur = off[a + 1] - off[a];
        remaining[a] = cur;
        if (cur > 1) {
            in_play[a] = 1;
            heap.push(Key(cur, a));
        } else if (cur == 1) {                                            /* :26-37 */
            if (nb[off[a]] != a) { *msg = "single-read groups should contain only the read itself"; return false; }
            out.push_back(std::vector<int32_t>(1, a));
        } else {
            *msg = "zero length read group";                              /* :38-40 */
            return false;
        }
    }
    while (!heap.empty()) {
        const Key top = heap.top();
        heap.pop();
        const int32_t v = top.second;
        if (!in_play[v] || remaining[v] != top.first) continue;           /* stale entry */
        if (top.first == 0) continue;                                     /* wiped-out node, :50-56 */
        in_play[v] = 0;                                                   /* pop_back of :73-74 */
        std::vector<int32_t> cluster;
        for (long long k = off[v]; k < off[v + 1]; ++k) {                 /* :76-100 */
            const int32_t w = nb[k];
            if (remaining[w] == 0) continue;
            cluster.push_back(w);
            remaining[w] = 0;
            for (long long k2 = off[w]; k2 < off[w + 1]; ++k2) {
                const int32_t x = nb[k2];
                if (remaining[x] > 0) {
                    --remaining[x];
                    if (in_play[x]) heap.push(Key(remaining[x], x));
                }
            }
        }
        out.push_back(std::move(cluster));
    }
    return true;
}

struct DevMem {
    void* p = nullptr;
    ~DevMem() { if (p) cudaFree(p); }
    bool alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1) == cudaSuccess; }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

}  // namespace

struct sarlacc_lists {
    Lists L;
};

extern "C" {

int64_t sarlacc_lists_count(const sarlacc_lists* r) { return r ? (int64_t)r->L.off.size() - 1 : 0; }
int64_t sarlacc_lists_values(const sarlacc_lists* r) { return r ? (int64_t)r->L.values.size() : 0; }
int sarlacc_lists_fetch(const sarlacc_lists* r, int64_t* off, int32_t* values) {
    if (!r) return sarlacc::set_error("list handle is NULL");
    std::memcpy(off, r->L.off.data(), sizeof(int64_t) * r->L.off.size());
    if (!r->L.values.empty()) std::memcpy(values, r->L.values.data(), sizeof(int32_t) * r->L.values.size());
    return 0;
}
void sarlacc_lists_free(sarlacc_lists* r) { delete r; }

/* mode 0: umi_group; mode 1: neighbour lists only (fast_levdist_test(sorted=TRUE) when there is one group) */
static sarlacc_lists* umi_run(int mode, const uint8_t* pool1, const int64_t* off1, int64_t n, int threshold1,
        const uint8_t* pool2, const int64_t* off2, int threshold2,
        const int64_t* group_off, const int32_t* members, int64_t ngroups, int device)
{
    if (!pool1 || !off1 || (ngroups > 0 && (!group_off || !members))) { sarlacc::set_error("UMI buffers must not be NULL"); return nullptr; }
    const bool two = pool2 != nullptr;
    if (two && !off2) { sarlacc::set_error("'umi1' and 'umi2' should have the same length"); return nullptr; }
    Seqs S1{pool1, off1}, S2{pool2, off2};
    const bool dbg = std::getenv("SARLACC_DEBUG_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    double t_sorted = t_begin, t_device = t_begin;
    /* entries = members of pre-groups with more than one read (src/umi_group.cpp:39-42), sorted per group */
    std::vector<int64_t> ent;            /* 0-based read index */
    std::vector<int32_t> local, gstart, gend;
    int64_t maxl1 = 0, maxl2 = 0;
    for (int64_t g = 0; g < ngroups; ++g) {
        const int64_t a = group_off[g], b = group_off[g + 1];
        for (int64_t k = a; k < b; ++k) {
            if (members[k] < 1 || members[k] > n) { sarlacc::set_error("pre-group index out of range"); return nullptr; }
        }
        if (b - a == 0 || (mode == 0 && b - a < 2)) continue;
        const size_t base = ent.size();
        std::vector<int32_t> ord((size_t)(b - a));
        const Seqs& So = two ? S2 : S1;   /* with two UMIs the UMI2 trie drives the order (umi_group.cpp:82-101) */
        bool short_keys = true;
        for (int64_t k = a; k < b && short_keys; ++k) short_keys = So.len(members[k] - 1) <= 21;
        if (short_keys) {
            /* up to 21 bases: the walk order is the order of 3-bit digits (rank + 1, 0 = end) packed into one word */
            std::vector<std::pair<uint64_t, int32_t> > keyed((size_t)(b - a));
            for (int64_t k = a; k < b; ++k) {
                const int64_t r = members[k] - 1;
                const uint8_t* p = So.ptr(r);
                const int64_t len = So.len(r);
                uint64_t key = 0;
                for (int64_t x = 0; x < 21; ++x) key = (key << 3) | (x < len ? (uint64_t)(trie_rank(p[x]) + 1) : 0u);
                keyed[(size_t)(k - a)] = std::make_pair(key, (int32_t)(k - a));
            }
            std::sort(keyed.begin(), keyed.end());
            for (size_t k = 0; k < ord.size(); ++k) ord[k] = keyed[k].second;
        } else {
            for (size_t k = 0; k < ord.size(); ++k) ord[k] = (int32_t)k;
            std::sort(ord.begin(), ord.end(), [&](int32_t x, int32_t y) {
                const int64_t rx = members[a + x] - 1, ry = members[a + y] - 1;
                if (So.less(rx, ry)) return true;
                if (So.less(ry, rx)) return false;
                return x < y;
            });
        }
        for (size_t k = 0; k < ord.size(); ++k) {
            const int64_t r = members[a + ord[k]] - 1;
            ent.push_back(r);
            local.push_back(ord[k]);
            gstart.push_back((int32_t)base);
            gend.push_back((int32_t)(base + ord.size()));
            maxl1 = std::max(maxl1, S1.len(r));
            if (two) maxl2 = std::max(maxl2, S2.len(r));
        }
    }
    const long long E = (long long)ent.size();
    /* Identical UMIs (PCR duplicates -- the reason UMIs exist) have identical neighbour lists, the reference's own
     * shortcut in sorted_trie::find (src/sorted_trie.cpp:253-257).  With one UMI the sorted order makes them adjacent,
     * so the device pass runs over the distinct sequences of each group and the lists are expanded on the host:
     * a matched sequence contributes all its reads, in index order -- exactly the walk order.  (With two UMIs the
     * order is (UMI2, index), which interleaves different UMI1s: no collapsing there.) */
    const bool collapse = !two && std::getenv("SARLACC_UMI_NO_COLLAPSE") == nullptr;
    std::vector<long long> ufirst;           /* entry index of every distinct sequence's first read (+ sentinel) */
    std::vector<int32_t> uniq_of((size_t)E);
    for (long long e = 0; e < E; ++e) {
        const bool fresh = e == gstart[(size_t)e] || !collapse || !S1.same(ent[(size_t)e], ent[(size_t)e - 1]);
        if (fresh) ufirst.push_back(e);
        uniq_of[(size_t)e] = (int32_t)ufirst.size() - 1;
    }
    const long long U = (long long)ufirst.size();
    ufirst.push_back(E);
    t_sorted = now();
    std::vector<long long> offs((size_t)E + 1, 0);
    std::vector<int32_t> nbrs;
    if (E > 0) {
        if (E > 0x7fffffffLL) { sarlacc::set_error("too many reads in multi-read pre-groups for one device pass"); return nullptr; }
        if (maxl1 > 64 || maxl2 > 64) { sarlacc::set_error("UMIs longer than 64 bases are outside the device kernel's envelope"); return nullptr; }
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { sarlacc::set_error("no CUDA device available (this library has no CPU fallback)"); return nullptr; }
        if (device < 0 || device >= ndev) { sarlacc::set_error("device index out of range"); return nullptr; }
        cudaSetDevice(device);
        const int W1 = std::max<int>(4, (int)((maxl1 + 3) & ~3LL)), W2 = two ? std::max<int>(4, (int)((maxl2 + 3) & ~3LL)) : 0;
        std::vector<uint8_t> h1((size_t)U * W1, 0), h2((size_t)U * std::max(W2, 1), 0), hl1((size_t)U), hl2((size_t)U, 0), hf((size_t)U);
        std::vector<int32_t> ugs((size_t)U), uge((size_t)U), uid((size_t)U);
        for (long long u = 0; u < U; ++u) {
            const long long e = ufirst[(size_t)u];
            const int64_t r = ent[(size_t)e];
            std::memcpy(&h1[(size_t)u * W1], S1.ptr(r), (size_t)S1.len(r));
            hl1[(size_t)u] = (uint8_t)S1.len(r);
            uint8_t f = S1.storable(r) ? 1 : 0;
            if (two) {
                std::memcpy(&h2[(size_t)u * W2], S2.ptr(r), (size_t)S2.len(r));
                hl2[(size_t)u] = (uint8_t)S2.len(r);
                if (S2.storable(r)) f |= 2;
            }
            hf[(size_t)u] = f;
            ugs[(size_t)u] = uniq_of[(size_t)gstart[(size_t)e]];
            uge[(size_t)u] = uniq_of[(size_t)gend[(size_t)e] - 1] + 1;
            uid[(size_t)u] = (int32_t)u;
        }
        DevMem d1, d2, dl1, dl2, df, dgs, dge, dloc, dcnt, doff, dnb, dfirst;
        bool ok = d1.alloc(h1.size()) && d2.alloc(h2.size()) && dl1.alloc(U) && dl2.alloc(U) && df.alloc(U) &&
                  dgs.alloc(sizeof(int32_t) * U) && dge.alloc(sizeof(int32_t) * U) && dloc.alloc(sizeof(int32_t) * U) &&
                  dcnt.alloc(sizeof(int32_t) * U) && doff.alloc(sizeof(long long) * (U + 1)) &&
                  dfirst.alloc(sizeof(int32_t) * (size_t)U * kUmiCap);
        if (!ok) { sarlacc::set_error("CUDA error: out of device memory in the UMI pass"); return nullptr; }
        cudaStream_t st = nullptr;
        cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        auto up = [&](DevMem& d, const void* h, size_t bytes) { cudaMemcpyAsync(d.p, h, bytes, cudaMemcpyHostToDevice, st); };
        up(d1, h1.data(), h1.size());
        if (two) up(d2, h2.data(), h2.size());
        up(dl1, hl1.data(), (size_t)U);
        up(dl2, hl2.data(), (size_t)U);
        up(df, hf.data(), (size_t)U);
        up(dgs, ugs.data(), sizeof(int32_t) * U);
        up(dge, uge.data(), sizeof(int32_t) * U);
        up(dloc, uid.data(), sizeof(int32_t) * U);
        UmiArgs A;
        std::memset(&A, 0, sizeof(A));
        A.seq1 = d1.as<uint8_t>();
        A.seq2 = two ? d2.as<uint8_t>() : nullptr;
        A.len1 = dl1.as<uint8_t>();
        A.len2 = dl2.as<uint8_t>();
        A.flags = df.as<uint8_t>();
        A.gstart = dgs.as<int32_t>();
        A.gend = dge.as<int32_t>();
        A.local = dloc.as<int32_t>();      /* the lists name distinct sequences; reads are filled in below */
        A.W1 = W1;
        A.W2 = W2;
        A.limit1 = 2 * threshold1;       /* limit *= MULT, src/sorted_trie.cpp:203 */
        A.limit2 = 2 * threshold2;
        A.E = U;
        A.count = dcnt.as<int32_t>();
        A.first = dfirst.as<int32_t>();
        bool launched = launch_umi(A, false, st);
        std::vector<int32_t> cnt((size_t)U);
        std::vector<long long> uoffs((size_t)U + 1, 0);
        std::vector<int32_t> unbrs;
        cudaMemcpyAsync(cnt.data(), dcnt.p, sizeof(int32_t) * U, cudaMemcpyDeviceToHost, st);
        cudaError_t ce = cudaStreamSynchronize(st);
        if (launched && ce == cudaSuccess) {
            for (long long u = 0; u < U; ++u) uoffs[(size_t)u + 1] = uoffs[(size_t)u] + cnt[(size_t)u];
            unbrs.resize((size_t)uoffs[(size_t)U]);
            if (!dnb.alloc(sizeof(int32_t) * unbrs.size())) { cudaStreamDestroy(st); sarlacc::set_error("CUDA error: out of device memory for the UMI neighbour lists"); return nullptr; }
            up(doff, uoffs.data(), sizeof(long long) * (U + 1));
            A.offset = doff.as<long long>();
            A.neighbors = dnb.as<int32_t>();
            launch_umi(A, true, st);
            if (!unbrs.empty()) cudaMemcpyAsync(unbrs.data(), dnb.p, sizeof(int32_t) * unbrs.size(), cudaMemcpyDeviceToHost, st);
            ce = cudaStreamSynchronize(st);
            sarlacc::count_launches(2);
        }
        cudaStreamDestroy(st);
        if (!launched) { sarlacc::set_error("no UMI kernel for this width"); return nullptr; }
        if (ce != cudaSuccess || (ce = cudaGetLastError()) != cudaSuccess) {
            sarlacc::set_error(std::string("CUDA error: ") + cudaGetErrorString(ce) + " in the UMI neighbour pass");
            return nullptr;
        }
        /* from distinct sequences back to reads: every read of a matched sequence, in entry (= index) order */
        for (long long e = 0; e < E; ++e) {
            const long long u = uniq_of[(size_t)e];
            long long c = 0;
            for (long long k = uoffs[(size_t)u]; k < uoffs[(size_t)u + 1]; ++k) {
                const long long v = unbrs[(size_t)k];
                c += ufirst[(size_t)v + 1] - ufirst[(size_t)v];
            }
            offs[(size_t)e + 1] = offs[(size_t)e] + c;
        }
        nbrs.resize((size_t)offs[(size_t)E]);
        for (long long e = 0; e < E; ++e) {
            const long long u = uniq_of[(size_t)e];
            long long at = offs[(size_t)e];
            for (long long k = uoffs[(size_t)u]; k < uoffs[(size_t)u + 1]; ++k) {
                const long long v = unbrs[(size_t)k];
                for (long long e2 = ufirst[(size_t)v]; e2 < ufirst[(size_t)v + 1]; ++e2) nbrs[(size_t)at++] = local[(size_t)e2];
            }
        }
    }

    t_device = now();
    /* back to pre-group order: lists indexed by the read's position in its group */
    std::unique_ptr<sarlacc_lists> res(new sarlacc_lists());
    long long e0 = 0;
    for (int64_t g = 0; g < ngroups; ++g) {
        const int64_t a = group_off[g], b = group_off[g + 1];
        const int cur = (int)(b - a);
        if (cur == 0) continue;                     /* an empty pre-group yields an empty list: nothing after unlist() */
        if (cur == 1 && mode == 0) {                /* src/umi_group.cpp:39-42 */
            res->L.push(std::vector<int32_t>(1, members[a]));
            continue;
        }
        /* CSR by local index */
        std::vector<long long> loff((size_t)cur + 1, 0);
        std::vector<long long> where((size_t)cur);
        for (int k = 0; k < cur; ++k) {
            const long long e = e0 + k;
            where[(size_t)local[(size_t)e]] = e;
        }
        for (int s = 0; s < cur; ++s) loff[(size_t)s + 1] = loff[(size_t)s] + (offs[(size_t)where[(size_t)s] + 1] - offs[(size_t)where[(size_t)s]]);
        std::vector<int32_t> lnb((size_t)loff[(size_t)cur]);
        for (int s = 0; s < cur; ++s) {
            const long long e = where[(size_t)s];
            std::copy(nbrs.begin() + offs[(size_t)e], nbrs.begin() + offs[(size_t)e + 1], lnb.begin() + loff[(size_t)s]);
        }
        if (mode == 1) {
            for (int s = 0; s < cur; ++s) {
                std::vector<int32_t> v(lnb.begin() + loff[(size_t)s], lnb.begin() + loff[(size_t)s + 1]);
                for (auto& x : v) x = members[a + x];
                res->L.push(v);
            }
        } else {
            std::vector<std::vector<int32_t> > clusters;
            const char* msg = nullptr;
            if (!cluster_lists(loff.data(), lnb.data(), cur, clusters, &msg)) { sarlacc::set_error(msg); return nullptr; }
            for (auto& c : clusters) {
                for (auto& x : c) x = members[a + x];                     /* src/umi_group.cpp:105-109 */
                res->L.push(c);
            }
        }
        e0 += cur;
    }
    if (dbg) std::fprintf(stderr, "[sarlacc] umi: sort %.1f ms, pack + device passes %.1f ms, lists + clustering %.1f ms (%lld reads in multi-read groups, %lld distinct sequences, %zu neighbours)\n",
                          (t_sorted - t_begin) * 1e3, (t_device - t_sorted) * 1e3, (now() - t_device) * 1e3, E, U, nbrs.size());
    return res.release();
}

/* cluster_umis_test (src/cluster_umis_test.cpp:8-29): the host clustering alone, 1-based links in, 1-based clusters out */
sarlacc_lists* sarlacc_cluster_umis(const int64_t* link_off, const int32_t* links, int64_t n)
{
    if (n < 0 || !link_off || (n > 0 && link_off[n] > 0 && !links)) { sarlacc::set_error("link lists must not be NULL"); return nullptr; }
    if (n > 0x7fffffffLL) { sarlacc::set_error("too many reads"); return nullptr; }
    std::vector<long long> off((size_t)n + 1);
    for (int64_t i = 0; i <= n; ++i) off[(size_t)i] = link_off[i] - link_off[0];
    std::vector<int32_t> nb((size_t)off[(size_t)n]);
    for (size_t k = 0; k < nb.size(); ++k) {
        const int32_t v = links[link_off[0] + (int64_t)k];
        if (v < 1 || v > n) { sarlacc::set_error("link index out of range"); return nullptr; }
        nb[k] = v - 1;
    }
    std::vector<std::vector<int32_t> > clusters;
    const char* msg = nullptr;
    if (!cluster_lists(off.data(), nb.data(), (int)n, clusters, &msg)) { sarlacc::set_error(msg); return nullptr; }
    std::unique_ptr<sarlacc_lists> res(new sarlacc_lists());
    for (auto& c : clusters) {
        for (auto& x : c) ++x;
        res->L.push(c);
    }
    return res.release();
}

sarlacc_lists* sarlacc_umi_group(const uint8_t* umi1_pool, const int64_t* umi1_off, int64_t n, int threshold1,
        const uint8_t* umi2_pool, const int64_t* umi2_off, int threshold2,
        const int64_t* group_off, const int32_t* group_members, int64_t ngroups, int device)
{
    return umi_run(0, umi1_pool, umi1_off, n, threshold1, umi2_pool, umi2_off, threshold2, group_off, group_members, ngroups, device);
}

sarlacc_lists* sarlacc_umi_neighbors(const uint8_t* umi1_pool, const int64_t* umi1_off, int64_t n, int threshold1,
        const uint8_t* umi2_pool, const int64_t* umi2_off, int threshold2,
        const int64_t* group_off, const int32_t* group_members, int64_t ngroups, int device)
{
    return umi_run(1, umi1_pool, umi1_off, n, threshold1, umi2_pool, umi2_off, threshold2, group_off, group_members, ngroups, device);
}

}
