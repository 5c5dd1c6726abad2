#include <cub/cub.cuh>

#include "prims.cuh"

namespace asrb {

static bool g_pool_ready = false;
void ensure_pool_configured() {
    if (g_pool_ready) return;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t thr = ~uint64_t(0);  // keep freed blocks cached in the pool
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    g_pool_ready = true;
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
