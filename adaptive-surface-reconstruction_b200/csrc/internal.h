// Internal (non-ABI) declarations shared by the .cu translation units.
#pragma once
#include <memory>
#include <vector>

#include "common.cuh"
#include "hash.cuh"

namespace asrb {

// Octree frame: cube, per-level voxel sizes, integer offset of the finest grid
// (reference: Octree::Octree, cpp/lib/octree.cpp:20-42).
struct Frame {
    float vs[kMaxLevel + 1];
    float ivs[kMaxLevel + 1];
    int off[3];
    float bb_min[3], bb_max[3];
};
Frame make_frame(const float* bb_min, const float* bb_max);

struct GridLevel {
    int64_t V = 0, E = 0;
    DevBuf<Key> keys;
    DevBuf<float> centers, sizes;
    DevBuf<int32_t> nidx;
    DevBuf<uint8_t> nslot;
    DevBuf<int64_t> nsplits;
    DevBuf<int32_t> uidx;  // up table (absent on the last level)
    DevBuf<uint8_t> uslot;
    bool has_up = false;
};

struct Octree {
    Frame frame;
    int device = 0;
    bool any = false;          // at least one point inside the bounding box
    bool root_separate = false;  // node list starts with the root key 1
    int64_t num_groups = 0;    // sibling groups (first-sibling keys), ascending
    int64_t num_nodes = 0;
    int64_t num_leaves = 0;
    DevBuf<Key> groups;
    KeyTable group_table;          // group key -> position in `groups`
    DevBuf<uint8_t> node_leaf;     // [num_nodes] 1 = leaf
    DevBuf<int64_t> node_rank;     // [num_nodes+1] exclusive scan of node_leaf
    DevBuf<Key> leaves;            // [num_leaves] ascending
    int balance_rounds = 0;
    // grid hierarchy (CreateGridsFromOctree), built lazily
    std::vector<std::unique_ptr<GridLevel>> grids;
    bool grids_all_info = false;
    // dual cells (CreateDualVertexIndices), built lazily
    int64_t num_duals = -1;
    DevBuf<uint8_t> dual_mask;     // [num_leaves] bit i = corner i emits a dual
    DevBuf<int64_t> dual_offset;   // [num_leaves+1]
    // asynchronous form (duals_begin / duals_fill): count and error flag travel to pinned slots behind events
    int64_t* dual_count_host = nullptr;
    int64_t* dual_err_host = nullptr;
    cudaEvent_t dual_count_event = nullptr, dual_err_event = nullptr;
    DevBuf<int> dual_err;
    bool dual_begun = false, dual_err_pending = false;
    ~Octree();
};

// octree.cu
void octree_build(Octree& t, const float* d_points, const float* d_radii, int64_t n, float radius_scale,
                  int max_depth, cudaStream_t s);
// grids.cu
void grids_build(Octree& t, int num_levels, bool all_info, cudaStream_t s);
void duals_begin(Octree& t, cudaStream_t s);   // queues the counting pass, no host synchronisation
void duals_count(Octree& t, cudaStream_t s);   // waits for it (or runs it)
void duals_check(Octree& t);                    // waits for duals_fill's error flag; throws like the reference
void duals_fill(Octree& t, int64_t* d_out, cudaStream_t s);

}  // namespace asrb
