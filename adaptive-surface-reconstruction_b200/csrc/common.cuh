// Shared device/host helpers for the asr_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace asrb {

// ---------------------------------------------------------------- errors
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
enum { kOk = 0, kInvalidArgument = 1, kRuntimeError = 2, kCudaError = 3 };

#define ASRB_CUDA(expr)                                                              \
    do {                                                                             \
        cudaError_t e__ = (expr);                                                    \
        if (e__ != cudaSuccess)                                                      \
            throw ::asrb::Error(::asrb::kCudaError,                                  \
                                std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)
// every hand-written kernel launch is followed by this; it also counts launches
extern std::atomic<long long> g_kernel_launches;
#define ASRB_CHECK_LAUNCH()                                             \
    do {                                                                \
        ::asrb::g_kernel_launches.fetch_add(1, std::memory_order_relaxed); \
        ASRB_CUDA(cudaGetLastError());                                  \
    } while (0)
#define ASRB_REQUIRE(cond, msg)                                                    \
    do {                                                                           \
        if (!(cond)) throw ::asrb::Error(::asrb::kInvalidArgument, std::string(msg)); \
    } while (0)

// ---------------------------------------------------------------- stream-ordered buffers
// Scratch and handle-owned storage come from the CUDA stream-ordered pool so the
// library never synchronises for an allocation and never touches torch's cache.
void ensure_pool_configured();

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaStream_t s = nullptr;
    DevBuf() {}
    DevBuf(size_t count, cudaStream_t stream) { alloc(count, stream); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), s(o.s) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) {
            release();
            p = o.p; n = o.n; s = o.s;
            o.p = nullptr; o.n = 0;
        }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t count, cudaStream_t stream) {
        release();
        ensure_pool_configured();
        n = count;
        s = stream;
        if (count) ASRB_CUDA(cudaMallocAsync((void**)&p, count * sizeof(T), stream));
    }
    void release() {
        if (p) cudaFreeAsync(p, s);
        p = nullptr;
        n = 0;
    }
    T* get() const { return p; }
    size_t size() const { return n; }
};

template <class T>
inline T d2h_scalar(const T* d, cudaStream_t s) {
    T v;
    ASRB_CUDA(cudaMemcpyAsync(&v, d, sizeof(T), cudaMemcpyDeviceToHost, s));
    ASRB_CUDA(cudaStreamSynchronize(s));
    return v;
}

inline unsigned grid_for(size_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

// ---------------------------------------------------------------- location codes
// key = morton(x, y, z) | 1 << 3*level   (reference: zindex.h:34, octreebase.h:59)
typedef unsigned long long Key;
constexpr int kMaxLevel = 21;
constexpr Key kNoKey = ~Key(0);

__host__ __device__ __forceinline__ Key spread3(Key v) {  // 21 bits -> every third bit
    v &= 0x1fffffULL;
    v = (v | v << 32) & 0x1f00000000ffffULL;
    v = (v | v << 16) & 0x1f0000ff0000ffULL;
    v = (v | v << 8) & 0x100f00f00f00f00fULL;
    v = (v | v << 4) & 0x10c30c30c30c30c3ULL;
    v = (v | v << 2) & 0x1249249249249249ULL;
    return v;
}
__host__ __device__ __forceinline__ unsigned compact3(Key v) {
    v &= 0x1249249249249249ULL;
    v = (v ^ (v >> 2)) & 0x10c30c30c30c30c3ULL;
    v = (v ^ (v >> 4)) & 0x100f00f00f00f00fULL;
    v = (v ^ (v >> 8)) & 0x1f0000ff0000ffULL;
    v = (v ^ (v >> 16)) & 0x1f00000000ffffULL;
    v = (v ^ (v >> 32)) & 0x1fffffULL;
    return (unsigned)v;
}
__host__ __device__ __forceinline__ Key morton3(unsigned x, unsigned y, unsigned z) {
    return spread3(x) | (spread3(y) << 1) | (spread3(z) << 2);
}

struct Cell {
    int x, y, z, lev;
};

__device__ __forceinline__ int key_level(Key k) { return (63 - __clzll((long long)k)) / 3; }

__device__ __forceinline__ bool cell_valid(int x, int y, int z, int lev) {
    const int n = 1 << lev;  // lev <= 21
    return lev >= 0 && lev <= kMaxLevel && x >= 0 && x < n && y >= 0 && y < n && z >= 0 && z < n;
}
// 0 == INVALID_KEY (octreebase.h:45)
__device__ __forceinline__ Key cell_key(int x, int y, int z, int lev) {
    if (!cell_valid(x, y, z, lev)) return 0;
    return morton3(x, y, z) | (Key(1) << (3 * lev));
}
__device__ __forceinline__ Cell key_cell(Key k) {
    Cell c;
    c.lev = key_level(k);
    k &= ~(Key(1) << (3 * c.lev));
    c.x = (int)compact3(k);
    c.y = (int)compact3(k >> 1);
    c.z = (int)compact3(k >> 2);
    return c;
}

// index of `k` in the ascending array `a[0..n)`, or -1
__device__ __forceinline__ long long find_key(const Key* __restrict__ a, long long n, Key k) {
    long long lo = 0, hi = n;
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < k) lo = mid + 1;
        else hi = mid;
    }
    return (lo < n && __ldg(a + lo) == k) ? lo : -1;
}
__device__ __forceinline__ long long lower_bound_key(const Key* __restrict__ a, long long n, Key k) {
    long long lo = 0, hi = n;
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < k) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

}  // namespace asrb
