// Dual contouring, vertex part: which dual cells the surface crosses and where.
//
// Replaces the first two passes of CreateTriangleMesh (reference
// cpp/lib/contouring.cpp:66-199).  Corner j of a dual is the leaf on the
// (-x if j&1, -y if j&2, -z if j&4) side of the octree vertex; an edge crosses
// when NOT both unsigned values exceed the threshold AND the signed values
// change sign strictly; the vertex is the mean of the linearly interpolated
// crossings, evaluated in double and rounded once.  One thread per dual;
// flag -> scan -> compute, output order = dual order.
#include "internal.h"
#include "prims.cuh"
#include "profile.cuh"

namespace asrb {

__constant__ int c_edges[12][2] = {{0, 1}, {1, 3}, {3, 2}, {2, 0}, {4, 5}, {5, 7},
                                   {7, 6}, {6, 4}, {0, 4}, {1, 5}, {3, 7}, {2, 6}};

__device__ __forceinline__ bool edge_crosses(const float2 a, const float2 b, float thr) {
    if (a.y > thr && b.y > thr) return false;
    return (a.x < 0.f && b.x > 0.f) || (a.x > 0.f && b.x < 0.f);
}

__global__ void __launch_bounds__(256)
contour_flag_kernel(const float2* __restrict__ values, const int64_t* __restrict__ duals, long long D, float thr,
                    uint8_t* __restrict__ flag) {
    const long long d = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (d >= D) return;
    float2 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __ldg(values + duals[8 * d + j]);
    bool hit = false;
#pragma unroll
    for (int e = 0; e < 12; ++e) hit |= edge_crosses(v[c_edges[e][0]], v[c_edges[e][1]], thr);
    flag[d] = hit ? 1 : 0;
}

__global__ void __launch_bounds__(256)
contour_vertex_kernel(const float2* __restrict__ values, const int64_t* __restrict__ duals, long long D, float thr,
                      const float* __restrict__ pos, const uint8_t* __restrict__ flag,
                      const int64_t* __restrict__ offset, float* __restrict__ vertices,
                      int64_t* __restrict__ vertex_dual) {
    const long long d = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (d >= D || !flag[d]) return;
    int64_t id[8];
    float2 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        id[j] = duals[8 * d + j];
        v[j] = __ldg(values + id[j]);
    }
    double acc[3] = {0.0, 0.0, 0.0};
    int cnt = 0;
#pragma unroll
    for (int e = 0; e < 12; ++e) {
        const int a = c_edges[e][0], b = c_edges[e][1];
        if (!edge_crosses(v[a], v[b], thr)) continue;
        const double v1 = (double)v[a].x, v2 = (double)v[b].x;
        double t = -v1 / (v2 - v1);
        if (!isfinite(t) || t < 0.0 || t > 1.0) t = 0.5;
#pragma unroll
        for (int k = 0; k < 3; ++k)
            acc[k] += __dadd_rn(__dmul_rn(1.0 - t, (double)pos[3 * id[a] + k]), __dmul_rn(t, (double)pos[3 * id[b] + k]));
        ++cnt;
    }
    const int64_t o = offset[d];
    vertices[3 * o + 0] = (float)(acc[0] / (double)cnt);
    vertices[3 * o + 1] = (float)(acc[1] / (double)cnt);
    vertices[3 * o + 2] = (float)(acc[2] / (double)cnt);
    if (vertex_dual) vertex_dual[o] = d;
}

void contour_count(const float* values, const int64_t* duals, int64_t D, float thr, uint8_t* flag, int64_t* offset,
                   int64_t* num_vertices, cudaStream_t s) {
    if (D) {
        ProfileScope prof("contour_flag", s);
        contour_flag_kernel<<<grid_for(D, 256), 256, 0, s>>>((const float2*)values, duals, D, thr, flag);
        ASRB_CHECK_LAUNCH();
    }
    exclusive_sum_u8_to_i64(flag, offset, (size_t)D, s);
    *num_vertices = d2h_scalar(offset + D, s);
}

void contour_fill(const float* values, const int64_t* duals, int64_t D, float thr, const float* pos,
                  const uint8_t* flag, const int64_t* offset, float* vertices, int64_t* vertex_dual, cudaStream_t s) {
    if (!D) return;
    ProfileScope prof("contour_vertex", s);
    contour_vertex_kernel<<<grid_for(D, 256), 256, 0, s>>>((const float2*)values, duals, D, thr, pos, flag, offset,
                                                           vertices, vertex_dual);
    ASRB_CHECK_LAUNCH();
}

}  // namespace asrb
