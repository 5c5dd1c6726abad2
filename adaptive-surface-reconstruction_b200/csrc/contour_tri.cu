// Dual contouring, triangle part: one polygon per primal edge the surface crosses.
//
// Replaces the second half of CreateTriangleMesh (reference cpp/lib/contouring.cpp:202-459).
// Vertex r of the mesh belongs to the r-th surface-crossing dual cell (contour.cu).  For every
// such dual and three of its twelve edges ({0,1}, {1,3}, {1,5}; the other nine belong to the
// neighbouring duals, :77-78) whose end nodes change sign, the crossing duals that contain
// BOTH end nodes are collected, ordered cyclically around the edge by walking across shared
// faces (starting in the direction given by the edge oriented from the smaller to the larger
// signed value, reversing once at an open end, :254-311), and emitted as 1 triangle (n = 3),
// 2 triangles split along the shorter diagonal (n = 4), or a fan around an extra centre
// vertex (n > 4); n < 3 emits nothing.
//
// The reference walks hash sets on one thread; here:
//   adjacency  node -> crossing duals (CSR; count with atomics, scan, fill, each list sorted so
//              the result does not depend on the atomic order)
//   count      thread = (vertex, edge): set intersection of the two nodes' lists -> n ->
//              #triangles, #extra vertices;  scans give every polygon its output position
//              (= the reference's emission order: vertex-major, edge-minor)
//   fill       same thread redoes the intersection, orders the duals and writes the triangles.
// The reference starts each cycle at an arbitrary (hash-order) dual, so triangle vertex order is
// defined up to rotation and the fan centre up to float summation order; tests compare
// rotation-normalised triangle sets.
#include "internal.h"
#include "prims.cuh"
#include "profile.cuh"

namespace asrb {

namespace {
constexpr int kMaxRing = 32;  // duals around one primal edge (reference: unbounded; > 32 -> error)

__constant__ int t_edges3[3][2] = {{0, 1}, {1, 3}, {1, 5}};
__constant__ int t_faces[6][4] = {{0, 1, 3, 2}, {4, 6, 7, 5}, {1, 5, 7, 3}, {2, 3, 7, 6}, {0, 2, 6, 4}, {0, 4, 5, 1}};

__device__ __forceinline__ bool edge_test(const float2* __restrict__ values, int64_t a, int64_t b, float thr) {
    if (a == b) return false;
    const float2 va = __ldg(values + a), vb = __ldg(values + b);
    if (va.y > thr && vb.y > thr) return false;
    return (va.x < 0.f && vb.x > 0.f) || (va.x > 0.f && vb.x < 0.f);
}

struct Face {  // SmallSet<size_t, 4>: ascending, unique
    int64_t v[4];
    int n;
};
__device__ __forceinline__ void face_insert(Face& f, int64_t x) {
    int i = 0;
    while (i < f.n && f.v[i] < x) ++i;
    if (i < f.n && f.v[i] == x) return;
    for (int j = f.n; j > i; --j) f.v[j] = f.v[j - 1];
    f.v[i] = x;
    ++f.n;
}
__device__ __forceinline__ bool face_equal(const Face& a, const Face& b) {
    if (a.n != b.n) return false;
    for (int i = 0; i < a.n; ++i)
        if (a.v[i] != b.v[i]) return false;
    return true;
}
__device__ __forceinline__ Face dual_face(const int64_t* __restrict__ d, int fi) {
    Face f;
    f.n = 0;
    for (int k = 0; k < 4; ++k) face_insert(f, d[t_faces[fi][k]]);
    return f;
}
// getDualFaceWithOrientedEdge (:234-252)
__device__ __forceinline__ Face face_with_oriented_edge(const int64_t* __restrict__ d, int64_t e0, int64_t e1) {
    for (int fi = 0; fi < 6; ++fi)
        for (int j = 0; j < 4; ++j)
            if (d[t_faces[fi][j]] == e0 && d[t_faces[fi][(j + 1) & 3]] == e1) {
                const Face f = dual_face(d, fi);
                if (f.n >= 3) return f;
            }
    Face none;
    none.n = 0;
    return none;
}
// DualHasFace (:219-232)
__device__ __forceinline__ bool dual_has_face(const int64_t* __restrict__ d, const Face& face) {
    for (int fi = 0; fi < 6; ++fi)
        if (face_equal(dual_face(d, fi), face)) return true;
    return false;
}

// crossing duals (vertex ids) that contain both nodes, ascending; returns n (or -1: more than kMaxRing)
__device__ __forceinline__ int common_duals(const int32_t* __restrict__ adj, const int64_t* __restrict__ adj_off, int64_t a,
                                            int64_t b, int32_t (&out)[kMaxRing]) {
    const int64_t a0 = adj_off[a], a1 = adj_off[a + 1], b0 = adj_off[b], b1 = adj_off[b + 1];
    int n = 0;
    int64_t j = b0;
    for (int64_t i = a0; i < a1; ++i) {  // both lists ascending: merge
        const int32_t x = adj[i];
        while (j < b1 && adj[j] < x) ++j;
        if (j < b1 && adj[j] == x) {
            if (n == kMaxRing) return -1;
            out[n++] = x;
        }
    }
    return n;
}
}  // namespace

// ------------------------------------------------------------------ node -> crossing duals
template <bool FILL>
__global__ void __launch_bounds__(256)
tri_adjacency_kernel(const int64_t* __restrict__ duals, const int64_t* __restrict__ vertex_dual, long long M,
                     int32_t* __restrict__ count, const int64_t* __restrict__ adj_off, int32_t* __restrict__ adj) {
    const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (r >= M) return;
    const int64_t* d = duals + 8 * vertex_dual[r];
    int64_t id[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) id[j] = d[j];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        bool first = true;
        for (int k = 0; k < j; ++k) first &= id[k] != id[j];
        if (!first) continue;
        const int pos = atomicAdd(count + id[j], 1);
        if (FILL) adj[adj_off[id[j]] + pos] = (int32_t)r;
    }
}

// insertion sort of every node's (short) list
__global__ void __launch_bounds__(256)
tri_adjacency_sort_kernel(const int64_t* __restrict__ adj_off, long long V, int32_t* __restrict__ adj) {
    const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (v >= V) return;
    const int64_t b = adj_off[v], e = adj_off[v + 1];
    for (int64_t i = b + 1; i < e; ++i) {
        const int32_t x = adj[i];
        int64_t j = i;
        while (j > b && adj[j - 1] > x) {
            adj[j] = adj[j - 1];
            --j;
        }
        adj[j] = x;
    }
}

// ------------------------------------------------------------------ polygons
__global__ void __launch_bounds__(256)
tri_count_kernel(const float2* __restrict__ values, const int64_t* __restrict__ duals,
                 const int64_t* __restrict__ vertex_dual, long long M, float thr, const int32_t* __restrict__ adj,
                 const int64_t* __restrict__ adj_off, int32_t* __restrict__ ntri, uint8_t* __restrict__ nextra,
                 int* __restrict__ error) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= 3 * M) return;
    const long long r = t / 3;
    const int ei = (int)(t - 3 * r);
    const int64_t* d = duals + 8 * vertex_dual[r];
    const int64_t a = d[t_edges3[ei][0]], b = d[t_edges3[ei][1]];
    int nt = 0, nx = 0;
    if (edge_test(values, a, b, thr)) {
        int32_t ring[kMaxRing];
        const int n = common_duals(adj, adj_off, a, b, ring);
        if (n < 0) *error = 1;
        else if (n == 3) nt = 1;
        else if (n == 4) nt = 2;
        else if (n > 4) {
            nt = n;
            nx = 1;
        }
    }
    ntri[t] = nt;
    nextra[t] = (uint8_t)nx;
}

__global__ void __launch_bounds__(128)
tri_fill_kernel(const float2* __restrict__ values, const int64_t* __restrict__ duals,
                const int64_t* __restrict__ vertex_dual, long long M, float thr, const int32_t* __restrict__ adj,
                const int64_t* __restrict__ adj_off, const int64_t* __restrict__ tri_off,
                const int64_t* __restrict__ extra_off, float* __restrict__ vertices, int32_t* __restrict__ triangles,
                int* __restrict__ error) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= 3 * M) return;
    if (tri_off[t + 1] == tri_off[t]) return;
    const long long r = t / 3;
    const int ei = (int)(t - 3 * r);
    const int64_t* d = duals + 8 * vertex_dual[r];
    int64_t e0 = d[t_edges3[ei][0]], e1 = d[t_edges3[ei][1]];
    int32_t rem[kMaxRing], ring[kMaxRing];
    int n = common_duals(adj, adj_off, e0, e1, rem);
    if (__ldg(values + e0).x > __ldg(values + e1).x) {  // orient the edge from the smaller signed value (:356-358)
        const int64_t tmp = e0;
        e0 = e1;
        e1 = tmp;
    }
    // sortDualsContainingEdge2 (:254-311)
    int ns = 0, nr = n - 1;
    ring[ns++] = rem[n - 1];
    bool reverse_again = false;
    for (int it = 0; it < n * n && nr > 0; ++it) {
        const Face face = face_with_oriented_edge(duals + 8 * vertex_dual[ring[ns - 1]], e0, e1);
        bool found = false;
        for (int k = 0; k < nr; ++k) {
            if (dual_has_face(duals + 8 * vertex_dual[rem[k]], face)) {
                ring[ns++] = rem[k];
                for (int q = k; q + 1 < nr; ++q) rem[q] = rem[q + 1];
                --nr;
                found = true;
                break;
            }
        }
        if (!found) {  // open end: continue from the other side
            for (int q = 0; q < ns / 2; ++q) {
                const int32_t tmp = ring[q];
                ring[q] = ring[ns - 1 - q];
                ring[ns - 1 - q] = tmp;
            }
            const int64_t tmp = e0;
            e0 = e1;
            e1 = tmp;
            reverse_again = !reverse_again;
        }
    }
    if (reverse_again)
        for (int q = 0; q < ns / 2; ++q) {
            const int32_t tmp = ring[q];
            ring[q] = ring[ns - 1 - q];
            ring[ns - 1 - q] = tmp;
        }
    int32_t* out = triangles + 3 * tri_off[t];
    if (ns != n) {  // the reference throws "cannot sort duals" (:365-369)
        *error = 2;
        for (int q = 0; q < 3 * (int)(tri_off[t + 1] - tri_off[t]); ++q) out[q] = 0;
        return;
    }
    if (n == 3) {
        out[0] = ring[0];
        out[1] = ring[1];
        out[2] = ring[2];
    } else if (n == 4) {
        float p[4][3];
        for (int q = 0; q < 4; ++q)
            for (int k = 0; k < 3; ++k) p[q][k] = vertices[3 * (size_t)ring[q] + k];
        auto sq = [&](int i, int j) {
            const float dx = __fsub_rn(p[i][0], p[j][0]), dy = __fsub_rn(p[i][1], p[j][1]), dz = __fsub_rn(p[i][2], p[j][2]);
            return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        };
        if (sq(0, 2) > sq(1, 3)) {
            out[0] = ring[0]; out[1] = ring[1]; out[2] = ring[3];
            out[3] = ring[1]; out[4] = ring[2]; out[5] = ring[3];
        } else {
            out[0] = ring[0]; out[1] = ring[1]; out[2] = ring[2];
            out[3] = ring[0]; out[4] = ring[2]; out[5] = ring[3];
        }
    } else {
        float c[3] = {0.f, 0.f, 0.f};
        for (int q = 0; q < n; ++q)
            for (int k = 0; k < 3; ++k) c[k] = __fadd_rn(c[k], vertices[3 * (size_t)ring[q] + k]);
        const int64_t ci = M + extra_off[t];
        for (int k = 0; k < 3; ++k) vertices[3 * ci + k] = c[k] / (float)n;
        for (int q = 0; q < n; ++q) {
            out[3 * q + 0] = ring[q];
            out[3 * q + 1] = ring[(q + 1) % n];
            out[3 * q + 2] = (int32_t)ci;
        }
    }
}

// ------------------------------------------------------------------ host
struct TriPlan {
    int64_t M = 0, D = 0, V = 0, T = 0, X = 0;
    float thr = 0.f;
    const float* values = nullptr;
    const int64_t* duals = nullptr;
    const int64_t* vertex_dual = nullptr;
    DevBuf<int64_t> adj_off, tri_off, extra_off;
    DevBuf<int32_t> adj;
};

TriPlan* contour_triangles_create(const float* values, const int64_t* duals, int64_t D, float thr,
                                  const int64_t* vertex_dual, int64_t M, int64_t V, int64_t* num_triangles,
                                  int64_t* num_extra, cudaStream_t s) {
    ASRB_REQUIRE(M < (int64_t(1) << 31) / 3, "contouring: too many vertices for int32 triangle indices");
    auto P = std::make_unique<TriPlan>();
    P->M = M;
    P->D = D;
    P->V = V;
    P->thr = thr;
    P->values = values;
    P->duals = duals;
    P->vertex_dual = vertex_dual;
    *num_triangles = 0;
    *num_extra = 0;
    if (M == 0) return P.release();
    ProfileScope prof("contour_triangles_count", s);
    DevBuf<int32_t> count((size_t)V, s);
    ASRB_CUDA(cudaMemsetAsync(count.get(), 0, (size_t)V * sizeof(int32_t), s));
    tri_adjacency_kernel<false><<<grid_for(M, 256), 256, 0, s>>>(duals, vertex_dual, M, count.get(), nullptr, nullptr);
    ASRB_CHECK_LAUNCH();
    P->adj_off.alloc((size_t)V + 1, s);
    exclusive_sum_i32_to_i64(count.get(), P->adj_off.get(), (size_t)V, s);
    const int64_t A = d2h_scalar(P->adj_off.get() + V, s);
    P->adj.alloc((size_t)A, s);
    ASRB_CUDA(cudaMemsetAsync(count.get(), 0, (size_t)V * sizeof(int32_t), s));
    tri_adjacency_kernel<true><<<grid_for(M, 256), 256, 0, s>>>(duals, vertex_dual, M, count.get(), P->adj_off.get(),
                                                               P->adj.get());
    ASRB_CHECK_LAUNCH();
    tri_adjacency_sort_kernel<<<grid_for(V, 256), 256, 0, s>>>(P->adj_off.get(), V, P->adj.get());
    ASRB_CHECK_LAUNCH();
    DevBuf<int32_t> ntri((size_t)3 * M, s);
    DevBuf<uint8_t> nextra((size_t)3 * M, s);
    DevBuf<int> err(1, s);
    ASRB_CUDA(cudaMemsetAsync(err.get(), 0, sizeof(int), s));
    tri_count_kernel<<<grid_for(3 * M, 256), 256, 0, s>>>((const float2*)values, duals, vertex_dual, M, thr, P->adj.get(),
                                                          P->adj_off.get(), ntri.get(), nextra.get(), err.get());
    ASRB_CHECK_LAUNCH();
    P->tri_off.alloc((size_t)3 * M + 1, s);
    P->extra_off.alloc((size_t)3 * M + 1, s);
    exclusive_sum_i32_to_i64(ntri.get(), P->tri_off.get(), (size_t)3 * M, s);
    exclusive_sum_u8_to_i64(nextra.get(), P->extra_off.get(), (size_t)3 * M, s);
    P->T = d2h_scalar(P->tri_off.get() + 3 * M, s);
    P->X = d2h_scalar(P->extra_off.get() + 3 * M, s);
    if (d2h_scalar(err.get(), s)) throw Error(kRuntimeError, "contouring: more than 32 dual cells around one edge");
    *num_triangles = P->T;
    *num_extra = P->X;
    return P.release();
}

void contour_triangles_fill(TriPlan& P, float* vertices, int32_t* triangles, cudaStream_t s) {
    if (P.M == 0 || P.T == 0) return;
    DevBuf<int> err(1, s);
    ASRB_CUDA(cudaMemsetAsync(err.get(), 0, sizeof(int), s));
    {
        ProfileScope prof("contour_triangles_fill", s);
        tri_fill_kernel<<<grid_for(3 * P.M, 128), 128, 0, s>>>((const float2*)P.values, P.duals, P.vertex_dual, P.M, P.thr,
                                                               P.adj.get(), P.adj_off.get(), P.tri_off.get(),
                                                               P.extra_off.get(), vertices, triangles, err.get());
        ASRB_CHECK_LAUNCH();
    }
    if (d2h_scalar(err.get(), s)) throw Error(kRuntimeError, "this should not happen: cannot sort duals");
}

void contour_triangles_destroy(TriPlan* P) { delete P; }

}  // namespace asrb
