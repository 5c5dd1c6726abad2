// invert_neighbors_list: CSR transpose of a neighbour list with an attribute
// carried along (Open3D-ML op; reference call site
// models/v0/net_definitions_torch.py:22-36,548-559 turns the fine->coarse "up"
// table into the coarse->fine table of the down convolutions).
// Stable radix sort of (target index, entry id); row boundaries of the result by
// binary search in the sorted targets; query row of each entry by binary search
// in the input row splits.
#include "internal.h"
#include "prims.cuh"

namespace asrb {

__global__ void __launch_bounds__(256)
invert_init_kernel(const int32_t* __restrict__ idx, long long E, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= E) return;
    keys[j] = (uint32_t)idx[j];
    vals[j] = (uint32_t)j;
}

__global__ void __launch_bounds__(256)
invert_splits_kernel(const uint32_t* __restrict__ sorted, long long E, long long num_points, int64_t* __restrict__ out) {
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p > num_points) return;
    long long lo = 0, hi = E;
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        if ((long long)sorted[mid] < p) lo = mid + 1;
        else hi = mid;
    }
    out[p] = lo;
}

__global__ void __launch_bounds__(256)
invert_fill_kernel(const uint32_t* __restrict__ perm, long long E, const int64_t* __restrict__ in_splits, long long Q,
                   const unsigned char* __restrict__ attrs, int attr_bytes, int32_t* __restrict__ out_idx,
                   unsigned char* __restrict__ out_attrs) {
    long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= E) return;
    const long long e = perm[j];
    // query row of entry e: last r with in_splits[r] <= e
    long long lo = 0, hi = Q;
    while (lo < hi) {
        long long mid = (lo + hi + 1) >> 1;
        if (in_splits[mid] <= e) lo = mid;
        else hi = mid - 1;
    }
    out_idx[j] = (int32_t)lo;
    for (int b = 0; b < attr_bytes; ++b) out_attrs[j * attr_bytes + b] = attrs[e * attr_bytes + b];
}

void invert_neighbors_list(int64_t num_points, const int32_t* idx, const int64_t* splits, int64_t Q, int64_t E,
                           const void* attrs, int attr_bytes, int32_t* out_idx, int64_t* out_splits, void* out_attrs,
                           cudaStream_t s) {
    DevBuf<uint32_t> keys((size_t)E, s), perm((size_t)E, s);
    if (E) {
        invert_init_kernel<<<grid_for(E, 256), 256, 0, s>>>(idx, E, keys.get(), perm.get());
        ASRB_CHECK_LAUNCH();
        int bits = 1;
        while (bits < 32 && (int64_t(1) << bits) < num_points) ++bits;
        sort_pairs_u32_u32(keys.get(), perm.get(), (size_t)E, s, bits);
    }
    invert_splits_kernel<<<grid_for(num_points + 1, 256), 256, 0, s>>>(keys.get(), E, num_points, out_splits);
    ASRB_CHECK_LAUNCH();
    if (E) {
        invert_fill_kernel<<<grid_for(E, 256), 256, 0, s>>>(perm.get(), E, splits, Q, (const unsigned char*)attrs,
                                                            attrs ? attr_bytes : 0, out_idx, (unsigned char*)out_attrs);
        ASRB_CHECK_LAUNCH();
    }
}

}  // namespace asrb
