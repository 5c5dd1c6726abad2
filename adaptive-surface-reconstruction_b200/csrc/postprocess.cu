// Connected components of a triangle mesh (vertices connected through shared triangles).
//
// Replaces ConnectedComponents (reference cpp/lib/postprocess.cpp:81-141), the depth-first
// walk behind RemoveConnectedComponents (:143-176) — SURVEY.md §8 row f-3.  The reference
// numbers components in the order of their smallest vertex; here every vertex gets that
// smallest vertex index as its label (min-label hooking over the triangle edges with pointer
// jumping, a few rounds to the fixed point), which orders the components identically.
#include "internal.h"
#include "profile.cuh"

namespace asrb {

__device__ __forceinline__ int32_t cc_root(const int32_t* __restrict__ parent, int32_t v) {
    int32_t p = parent[v];
    while (p != v) {
        v = p;
        p = parent[v];
    }
    return v;
}

__global__ void __launch_bounds__(256)
cc_init_kernel(int32_t* __restrict__ parent, long long V) {
    const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (v < V) parent[v] = (int32_t)v;
}

__global__ void __launch_bounds__(256)
cc_hook_kernel(const int32_t* __restrict__ tri, long long T, long long V, int32_t* __restrict__ parent,
               int* __restrict__ changed, int* __restrict__ error) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= T) return;
    int32_t r[3];
    for (int k = 0; k < 3; ++k) {
        const int32_t v = tri[3 * t + k];
        if (v < 0 || v >= V) {  // the reference throws std::out_of_range -> RuntimeError (:52-58)
            *error = 1;
            return;
        }
        r[k] = cc_root(parent, v);
    }
    const int32_t m = min(r[0], min(r[1], r[2]));
    for (int k = 0; k < 3; ++k)
        if (r[k] != m) {
            atomicMin(parent + r[k], m);
            *changed = 1;
        }
}

__global__ void __launch_bounds__(256)
cc_compress_kernel(int32_t* __restrict__ parent, long long V) {
    const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (v < V) parent[v] = cc_root(parent, (int32_t)v);
}

__global__ void __launch_bounds__(256)
cc_widen_kernel(const int32_t* __restrict__ parent, long long V, int64_t* __restrict__ label,
                int64_t* __restrict__ size) {
    const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (v >= V) return;
    const int32_t r = parent[v];
    label[v] = r;
    if (size) atomicAdd(reinterpret_cast<unsigned long long*>(size + r), 1ULL);
}

// label[v] = smallest vertex index of v's component; size[r] = number of vertices of the
// component whose label is r (0 elsewhere; may be null)
void mesh_components(const int32_t* triangles, int64_t T, int64_t V, int64_t* label, int64_t* size, cudaStream_t s) {
    ASRB_REQUIRE(V < (int64_t(1) << 31), "mesh_components: too many vertices");
    if (V == 0) return;
    ProfileScope prof("mesh_components", s);
    DevBuf<int32_t> parent((size_t)V, s);
    DevBuf<int> flags(2, s);
    cc_init_kernel<<<grid_for(V, 256), 256, 0, s>>>(parent.get(), V);
    ASRB_CHECK_LAUNCH();
    ASRB_CUDA(cudaMemsetAsync(flags.get(), 0, 2 * sizeof(int), s));
    for (int round = 0; T > 0 && round < 64; ++round) {
        cc_hook_kernel<<<grid_for(T, 256), 256, 0, s>>>(triangles, T, V, parent.get(), flags.get(), flags.get() + 1);
        ASRB_CHECK_LAUNCH();
        cc_compress_kernel<<<grid_for(V, 256), 256, 0, s>>>(parent.get(), V);
        ASRB_CHECK_LAUNCH();
        int h[2];
        ASRB_CUDA(cudaMemcpyAsync(h, flags.get(), sizeof(h), cudaMemcpyDeviceToHost, s));
        ASRB_CUDA(cudaStreamSynchronize(s));
        if (h[1]) throw Error(kRuntimeError, "triangle index out of range");
        if (!h[0]) break;
        ASRB_CUDA(cudaMemsetAsync(flags.get(), 0, sizeof(int), s));
    }
    if (size) ASRB_CUDA(cudaMemsetAsync(size, 0, (size_t)V * sizeof(int64_t), s));
    cc_widen_kernel<<<grid_for(V, 256), 256, 0, s>>>(parent.get(), V, label, size);
    ASRB_CHECK_LAUNCH();
}

}  // namespace asrb
