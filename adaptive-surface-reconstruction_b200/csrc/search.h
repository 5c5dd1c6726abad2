#pragma once
#include "common.cuh"
#include "hash.cuh"

namespace asrb {

struct Float4Pod {
    float x, y, z, w;
};

struct Search {
    int64_t n = 0, nq = 0, num_pairs = 0;
    const float* queries = nullptr;  // borrowed until search_fill
    const float* radii = nullptr;
    float frame_origin[3] = {0, 0, 0};
    float frame_inv_h = 1.f;
    DevBuf<Key> codes;       // sorted Morton codes of the points
    DevBuf<Float4Pod> spts;  // points in sorted order, w = original index bits
    KeyTable cells;          // (cell code | level marker) -> begin | end << 32 in the sorted points
    DevBuf<int64_t> splits;  // [nq+1]
    bool want_fill = true;   // keep the first hits of every query from the count pass (search_fill follows)
    DevBuf<unsigned long long> stash;
};

void search_prepare(Search& S, const float* d_points, int64_t n, const float* d_queries, const float* d_radii,
                    int64_t nq, const float* h_frame, cudaStream_t s);
void search_fill(Search& S, int32_t* d_idx, float* d_d2, int64_t* d_splits, cudaStream_t s);
void scale_compat(const float* d_sizes, const float* d_radii, const int32_t* d_idx, const int64_t* d_splits,
                  int64_t nq, float* d_out, cudaStream_t s);

// kNN over the points themselves (reference KDTree, nsearch.cpp:30-105); S holds the sorted points
void knn_build(Search& S, const float* d_points, int64_t n, cudaStream_t s);
void knn_radius(const Search& S, int k, float* d_out, cudaStream_t s);
void knn_inlier(const Search& S, const float* d_radii, float fraction, int k, int outlier_threshold, uint8_t* d_out,
                cudaStream_t s);
void radius_neighbor_counts(const float* d_points, int64_t n, const float* d_radii, int32_t* d_out, cudaStream_t s);

}  // namespace asrb
