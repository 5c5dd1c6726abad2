#include <cub/cub.cuh>

#include "prims.cuh"

#include <cstdlib>
#include <mutex>
#include <vector>

namespace asrb {

// The library's scratch and handle-owned storage come from the device's default stream-ordered pool.  Two settings,
// once per device: (i) freed blocks stay in the pool (release threshold = max), (ii) the pool is grown ONCE to
// ASR_POOL_PREALLOC_GB (default 16) in one piece.  Without (ii) the pool holds just the high-water mark of a pass
// (~4 GB for 10 M points); the GB-sized transient buffers of the search and of the conv plans then regularly fail
// to find a contiguous free range and the driver re-maps physical memory inside the pool — measured in round 2 as
// sporadic host stalls of 30-600 ms inside cudaMallocAsync.
static std::mutex g_pool_mutex;
static bool g_pool_ready[64] = {};
void ensure_pool_configured() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return;
    if (g_pool_ready[dev]) return;
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (g_pool_ready[dev]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t thr = ~uint64_t(0);  // keep freed blocks cached in the pool
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        double gb = 16.0;
        if (const char* e = getenv("ASR_POOL_PREALLOC_GB")) gb = atof(e);
        size_t free_b = 0, total_b = 0;
        if (gb > 0 && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            size_t want = (size_t)(gb * (double)(1ull << 30));
            want = std::min(want, free_b / 4);  // never more than a quarter of what is free
            void* p = nullptr;
            if (want > (64u << 20) && cudaMallocAsync(&p, want, (cudaStream_t)0) == cudaSuccess) {
                cudaFreeAsync(p, (cudaStream_t)0);
                cudaStreamSynchronize((cudaStream_t)0);
            } else {
                cudaGetLastError();
            }
        }
    }
    g_pool_ready[dev] = true;
}

// pinned int64 slots for small asynchronous device-to-host results (cudaMallocHost / cudaFreeHost synchronise the
// device, so the slots are pooled and never freed)
static std::mutex g_slot_mutex;
static std::vector<int64_t*> g_slot_free;
int64_t* pinned_slot_acquire() {
    std::lock_guard<std::mutex> lock(g_slot_mutex);
    if (g_slot_free.empty()) {
        int64_t* page = nullptr;
        ASRB_CUDA(cudaMallocHost((void**)&page, 128 * sizeof(int64_t)));
        for (int i = 0; i < 128; ++i) g_slot_free.push_back(page + i);
    }
    int64_t* p = g_slot_free.back();
    g_slot_free.pop_back();
    return p;
}
void pinned_slot_release(int64_t* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lock(g_slot_mutex);
    g_slot_free.push_back(p);
}

template <class K>
static void sort_keys_impl(K* d_keys, size_t n, cudaStream_t s, int end_bit) {
    if (n < 2) return;
    DevBuf<K> alt(n, s);
    cub::DoubleBuffer<K> db(d_keys, alt.get());
    size_t bytes = 0;
    ASRB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, bytes, db, (int64_t)n, 0, end_bit, s));
    DevBuf<char> tmp(bytes, s);
    ASRB_CUDA(cub::DeviceRadixSort::SortKeys(tmp.get(), bytes, db, (int64_t)n, 0, end_bit, s));
    if (db.Current() != d_keys)
        ASRB_CUDA(cudaMemcpyAsync(d_keys, db.Current(), n * sizeof(K), cudaMemcpyDeviceToDevice, s));
}

template <class K, class V>
static void sort_pairs_impl(K* d_keys, V* d_vals, size_t n, cudaStream_t s, int end_bit) {
    if (n < 2) return;
    DevBuf<K> altk(n, s);
    DevBuf<V> altv(n, s);
    cub::DoubleBuffer<K> dk(d_keys, altk.get());
    cub::DoubleBuffer<V> dv(d_vals, altv.get());
    size_t bytes = 0;
    ASRB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, dk, dv, (int64_t)n, 0, end_bit, s));
    DevBuf<char> tmp(bytes, s);
    ASRB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.get(), bytes, dk, dv, (int64_t)n, 0, end_bit, s));
    if (dk.Current() != d_keys)
        ASRB_CUDA(cudaMemcpyAsync(d_keys, dk.Current(), n * sizeof(K), cudaMemcpyDeviceToDevice, s));
    if (dv.Current() != d_vals)
        ASRB_CUDA(cudaMemcpyAsync(d_vals, dv.Current(), n * sizeof(V), cudaMemcpyDeviceToDevice, s));
}

void sort_keys_u64(Key* d_keys, size_t n, cudaStream_t s, int end_bit) { sort_keys_impl(d_keys, n, s, end_bit); }
void sort_pairs_u64_u32(Key* k, uint32_t* v, size_t n, cudaStream_t s, int end_bit) { sort_pairs_impl(k, v, n, s, end_bit); }
void sort_pairs_u64_u64(Key* k, unsigned long long* v, size_t n, cudaStream_t s, int end_bit) { sort_pairs_impl(k, v, n, s, end_bit); }
void sort_pairs_u32_u32(uint32_t* k, uint32_t* v, size_t n, cudaStream_t s, int end_bit) { sort_pairs_impl(k, v, n, s, end_bit); }
void sort_pairs_u8_u32(uint8_t* k, uint32_t* v, size_t n, cudaStream_t s, int end_bit) { sort_pairs_impl(k, v, n, s, end_bit); }

size_t unique_u64(Key* d_keys, size_t n, cudaStream_t s) {
    if (n < 2) return n;
    DevBuf<Key> out(n, s);
    DevBuf<int64_t> cnt(1, s);
    size_t bytes = 0;
    ASRB_CUDA(cub::DeviceSelect::Unique(nullptr, bytes, d_keys, out.get(), cnt.get(), (int64_t)n, s));
    DevBuf<char> tmp(bytes, s);
    ASRB_CUDA(cub::DeviceSelect::Unique(tmp.get(), bytes, d_keys, out.get(), cnt.get(), (int64_t)n, s));
    int64_t m = d2h_scalar(cnt.get(), s);
    ASRB_CUDA(cudaMemcpyAsync(d_keys, out.get(), (size_t)m * sizeof(Key), cudaMemcpyDeviceToDevice, s));
    return (size_t)m;
}

template <class T>
__global__ void widen_kernel(const T* __restrict__ in, int64_t* __restrict__ out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int64_t)in[i];
    else if (i == n) out[i] = 0;
}

template <class T>
static void exclusive_sum_impl(const T* d_in, int64_t* d_out, size_t n, cudaStream_t s) {
    // widen into the output (plus one trailing zero), then scan n+1 items in
    // place so that out[n] holds the total
    widen_kernel<T><<<grid_for(n + 1, 256), 256, 0, s>>>(d_in, d_out, n);
    ASRB_CHECK_LAUNCH();
    size_t bytes = 0;
    ASRB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, d_out, d_out, (int64_t)(n + 1), s));
    DevBuf<char> tmp(bytes, s);
    ASRB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.get(), bytes, d_out, d_out, (int64_t)(n + 1), s));
}

void exclusive_sum_i32_to_i64(const int32_t* d_in, int64_t* d_out, size_t n, cudaStream_t s) { exclusive_sum_impl(d_in, d_out, n, s); }
void exclusive_sum_u8_to_i64(const uint8_t* d_in, int64_t* d_out, size_t n, cudaStream_t s) { exclusive_sum_impl(d_in, d_out, n, s); }

}  // namespace asrb
