/* .compute_threshold (/root/reference/R/getAdaptorThresholds.R:94-103) on the device.
 *
 *     real <- sort(real); scrambled <- sort(scrambled)
 *     fdr <- (length(scrambled) - findInterval(real, scrambled)) / (length(real) - seq_along(real))
 *     real[min(which(fdr <= error))]
 *
 * The one step of adaptorAlign + getAdaptorThresholds that looks at all reads at once: with 50 M reads per
 * adaptor the two sorts are what costs, so they run on the device (CUB radix sort -- library code, like the R sort it
 * replaces; the hot path is the alignment) and the FDR scan is one pass of binary searches with a min-index reduction.
 * IEEE semantics as in R: integer counts divided as doubles (correctly rounded division), the last term divides by zero
 * (Inf or NaN, never <= error unless error is Inf), an empty which() gives NA -> NaN here.
 */
#include "sarlacc_b200.h"
#include "kernels.h"

#include <cub/device/device_radix_sort.cuh>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <mutex>
#include <string>

namespace sarlacc {
int set_error(const std::string& msg);
void count_launches(int n);
void threshold_trim();
}

namespace {

/* findInterval(x, v) for sorted v = number of elements <= x. */
__device__ __forceinline__ long long count_le(const double* v, long long n, double x) {
    long long lo = 0, hi = n;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (v[mid] <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) fdr_first_index(const double* real, long long nr, const double* scr, long long ns, double error,
                                                      unsigned long long* first)
{
    long long best = nr;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < nr; k += (long long)gridDim.x * blockDim.x) {
        const double num = (double)(ns - count_le(scr, ns, real[k]));
        const double den = (double)(nr - (k + 1));
        const double fdr = __ddiv_rn(num, den);
        if (fdr <= error) { best = k; break; }       /* k ascends per thread: its first hit is its smallest */
    }
    /* min over the block, then over the grid */
    for (int o = 16; o > 0; o >>= 1) {
        const long long other = __shfl_down_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
    }
    if ((threadIdx.x & 31) == 0 && best < nr) atomicMin(first, (unsigned long long)best);
}

/* findInterval(x, v, left.open=TRUE) = number of elements < x. */
__device__ __forceinline__ long long count_lt(const double* v, long long n, double x) {
    long long lo = 0, hi = n;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (v[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) tied_overlap_sum(const double* real, long long nr, const double* fake, long long nf, unsigned long long* sum)
{
    unsigned long long mine = 0;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < nr; k += (long long)gridDim.x * blockDim.x) {
        const double x = real[k];
        mine += (unsigned long long)(count_le(fake, nf, x) + count_lt(fake, nf, x));
    }
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(sum, mine);
}

/* Work buffers are kept between calls (per device, grow-only; handed back by sarlacc_trim_device_memory): cudaMalloc and
 * cudaFree of a few hundred MB cost more than sorting them. */
struct Cache {
    std::mutex m;
    struct Slot { void* p = nullptr; size_t cap = 0; int dev = -1; };
    Slot slots[8][6];
    void* get(int dev, int k, size_t bytes, cudaError_t* err) {
        Slot& s = slots[dev & 7][k];
        *err = cudaSuccess;
        if (s.p && s.dev == dev && s.cap >= bytes) return s.p;
        if (s.p) cudaFree(s.p);
        s.p = nullptr;
        s.cap = 0;
        const size_t want = bytes + bytes / 8 + 256;
        *err = cudaMalloc(&s.p, want);
        if (*err != cudaSuccess) { s.p = nullptr; return nullptr; }
        s.cap = want;
        s.dev = dev;
        return s.p;
    }
    void trim() {
        std::lock_guard<std::mutex> lock(m);
        for (auto& d : slots) for (auto& s : d) {
            if (s.p) { cudaSetDevice(s.dev); cudaFree(s.p); }
            s = Slot();
        }
    }
};
Cache& cache() {
    static Cache* c = new Cache();      /* never destroyed: no CUDA calls during process teardown */
    return *c;
}

struct Tmp {          /* one cached buffer, borrowed for the duration of a call (the cache mutex is held by the caller) */
    void* p = nullptr;
    int dev = 0, slot = 0;
    cudaError_t alloc(size_t bytes) {
        cudaError_t e;
        p = cache().get(dev, slot, bytes ? bytes : 1, &e);
        return e;
    }
};

}  // namespace

void sarlacc::threshold_trim() { cache().trim(); }

extern "C" int sarlacc_compute_threshold(const double* real, int64_t nreal, const double* scrambled, int64_t nscr, double error,
                                         int device, double* threshold)
{
    if (!threshold) return sarlacc::set_error("threshold must not be NULL");
    if (nreal < 0 || nscr < 0 || (nreal > 0 && !real) || (nscr > 0 && !scrambled)) return sarlacc::set_error("score vectors must not be NULL");
    *threshold = std::numeric_limits<double>::quiet_NaN();       /* real[min(integer(0))] -> NA */
    if (nreal == 0) return 0;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
        cudaGetLastError();
        return sarlacc::set_error("sarlacc_b200 requires a CUDA device (no CPU fallback exists): no device found");
    }
#define TH_CHECK(expr)                                                                                      \
    do {                                                                                                    \
        cudaError_t e_ = (expr);                                                                            \
        if (e_ != cudaSuccess) return sarlacc::set_error(std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #expr); \
    } while (0)
    TH_CHECK(cudaSetDevice(device));
    const bool dbg = std::getenv("SARLACC_DEBUG_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    std::lock_guard<std::mutex> lock(cache().m);
    Tmp r_in{nullptr, device, 0}, r_out{nullptr, device, 1}, s_in{nullptr, device, 2}, s_out{nullptr, device, 3}, work{nullptr, device, 4}, first{nullptr, device, 5};
    const size_t rb = sizeof(double) * (size_t)nreal, sb = sizeof(double) * (size_t)nscr;
    TH_CHECK(r_in.alloc(rb));
    TH_CHECK(r_out.alloc(rb));
    TH_CHECK(s_in.alloc(sb));
    TH_CHECK(s_out.alloc(sb));
    TH_CHECK(first.alloc(sizeof(unsigned long long)));
    /* cudaMemcpyDefault: the vectors may live on the host or on a device (e.g. gathered from all ranks) */
    const double t1 = now();
    TH_CHECK(cudaMemcpy(r_in.p, real, rb, cudaMemcpyDefault));
    if (nscr > 0) TH_CHECK(cudaMemcpy(s_in.p, scrambled, sb, cudaMemcpyDefault));
    size_t wbytes = 0, wbytes2 = 0;
    TH_CHECK(cub::DeviceRadixSort::SortKeys(nullptr, wbytes, (const double*)r_in.p, (double*)r_out.p, (long long)nreal));
    TH_CHECK(cub::DeviceRadixSort::SortKeys(nullptr, wbytes2, (const double*)s_in.p, (double*)s_out.p, (long long)nscr));
    if (wbytes2 > wbytes) wbytes = wbytes2;
    TH_CHECK(work.alloc(wbytes));
    TH_CHECK(cub::DeviceRadixSort::SortKeys(work.p, wbytes, (const double*)r_in.p, (double*)r_out.p, (long long)nreal));
    if (nscr > 0) TH_CHECK(cub::DeviceRadixSort::SortKeys(work.p, wbytes, (const double*)s_in.p, (double*)s_out.p, (long long)nscr));
    if (dbg) cudaDeviceSynchronize();
    const double t2 = now();
    const unsigned long long none = ~0ULL;
    TH_CHECK(cudaMemcpy(first.p, &none, sizeof(none), cudaMemcpyHostToDevice));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    long long grid = (nreal + 255) / 256;
    if (grid > (long long)sms * 8) grid = (long long)sms * 8;
    fdr_first_index<<<(int)grid, 256>>>((const double*)r_out.p, (long long)nreal, (const double*)s_out.p, (long long)nscr, error,
                                        (unsigned long long*)first.p);
    sarlacc::count_launches(1);
    TH_CHECK(cudaGetLastError());
    unsigned long long k = none;
    TH_CHECK(cudaMemcpy(&k, first.p, sizeof(k), cudaMemcpyDeviceToHost));
    if (k != none) TH_CHECK(cudaMemcpy(threshold, (const double*)r_out.p + k, sizeof(double), cudaMemcpyDeviceToHost));
    if (dbg) std::fprintf(stderr, "[sarlacc] compute_threshold: %lld real, %lld scrambled scores: alloc %.2f ms, copy + sorts %.2f ms, scan %.2f ms (first index %lld)\n",
                          (long long)nreal, (long long)nscr, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (now() - t2) * 1e3, k == none ? -1LL : (long long)k);
#undef TH_CHECK
    return 0;
}

extern "C" int sarlacc_tied_overlap(const double* real, int64_t nreal, const double* fake, int64_t nfake, int device, double* overlap)
{
    if (!overlap) return sarlacc::set_error("overlap must not be NULL");
    if (nreal < 0 || nfake < 0 || (nreal > 0 && !real) || (nfake > 0 && !fake)) return sarlacc::set_error("score vectors must not be NULL");
    *overlap = std::numeric_limits<double>::quiet_NaN();      /* 0 / 0 in R */
    if (nreal == 0 || nfake == 0) return 0;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
        cudaGetLastError();
        return sarlacc::set_error("sarlacc_b200 requires a CUDA device (no CPU fallback exists): no device found");
    }
#define TH_CHECK(expr)                                                                                      \
    do {                                                                                                    \
        cudaError_t e_ = (expr);                                                                            \
        if (e_ != cudaSuccess) return sarlacc::set_error(std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #expr); \
    } while (0)
    TH_CHECK(cudaSetDevice(device));
    std::lock_guard<std::mutex> lock(cache().m);
    Tmp r_in{nullptr, device, 0}, f_in{nullptr, device, 2}, f_out{nullptr, device, 3}, work{nullptr, device, 4}, sum{nullptr, device, 5};
    const size_t rb = sizeof(double) * (size_t)nreal, fb = sizeof(double) * (size_t)nfake;
    TH_CHECK(r_in.alloc(rb));
    TH_CHECK(f_in.alloc(fb));
    TH_CHECK(f_out.alloc(fb));
    TH_CHECK(sum.alloc(sizeof(unsigned long long)));
    TH_CHECK(cudaMemcpy(r_in.p, real, rb, cudaMemcpyDefault));
    TH_CHECK(cudaMemcpy(f_in.p, fake, fb, cudaMemcpyDefault));
    size_t wbytes = 0;
    TH_CHECK(cub::DeviceRadixSort::SortKeys(nullptr, wbytes, (const double*)f_in.p, (double*)f_out.p, (long long)nfake));
    TH_CHECK(work.alloc(wbytes));
    TH_CHECK(cub::DeviceRadixSort::SortKeys(work.p, wbytes, (const double*)f_in.p, (double*)f_out.p, (long long)nfake));
    TH_CHECK(cudaMemset(sum.p, 0, sizeof(unsigned long long)));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    long long grid = (nreal + 255) / 256;
    if (grid > (long long)sms * 8) grid = (long long)sms * 8;
    tied_overlap_sum<<<(int)grid, 256>>>((const double*)r_in.p, (long long)nreal, (const double*)f_out.p, (long long)nfake, (unsigned long long*)sum.p);
    sarlacc::count_launches(1);
    TH_CHECK(cudaGetLastError());
    unsigned long long total = 0;
    TH_CHECK(cudaMemcpy(&total, sum.p, sizeof(total), cudaMemcpyDeviceToHost));
    /* every term (upper + lower) / 2 is a multiple of 0.5 and the sum stays far below 2^53, so R's double sum is exact */
    *overlap = ((double)total / 2.0) / ((double)nreal * (double)nfake);
#undef TH_CHECK
    return 0;
}
