// TEST INFRASTRUCTURE ONLY.  C-ABI wrapper around the *unmodified* reference
// geometry sources (compiled from /root/reference/cpp/lib by oracle/Makefile
// into oracle/_ref/libasr_ref.so).  Nothing in the product path links or loads
// this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may.
//
// Wrapped reference entry points:
//   CreateOctreeFromPoints      cpp/lib/octree.cpp:230
//   CreateGridsFromOctree       cpp/lib/grid.cpp:245
//   CreateDualVertexIndices     cpp/lib/grid.cpp:450
//   CreateTriangleMesh          cpp/lib/contouring.cpp:29
//   RemoveConnectedComponents   cpp/lib/postprocess.cpp:141
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "contouring.h"
#include "grid.h"
#include "octree.h"
#include "postprocess.h"

namespace {
thread_local std::string g_err;

struct RefTree {
    std::shared_ptr<asr::Octree> tree;
    std::vector<uint64_t> node_keys;  // all hash-map keys, sorted (for tests)
};
struct RefGrids {
    std::vector<std::shared_ptr<ASRGrid>> grids;
};
struct RefMesh {
    std::vector<float> vertices;
    std::vector<int32_t> triangles;
};
struct RefDuals {
    std::vector<size_t> idx;
};
}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

void* ref_octree_create(const float* points, uint64_t n, const float* radii,
                        const float* bb_min, const float* bb_max,
                        float radius_scale, int grow_steps, int max_depth) {
    try {
        auto* t = new RefTree;
        asr::Vec3f mn(bb_min[0], bb_min[1], bb_min[2]);
        asr::Vec3f mx(bb_max[0], bb_max[1], bb_max[2]);
        t->tree = asr::CreateOctreeFromPoints(points, n, radii, mn, mx,
                                              radius_scale, grow_steps,
                                              max_depth);
        return t;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
void ref_octree_free(void* h) { delete static_cast<RefTree*>(h); }

uint64_t ref_octree_num_leaves(void* h) {
    return static_cast<RefTree*>(h)->tree->leaves.size();
}
void ref_octree_leaves(void* h, uint64_t* out) {
    auto& l = static_cast<RefTree*>(h)->tree->leaves;
    std::memcpy(out, l.data(), l.size() * sizeof(uint64_t));
}
uint64_t ref_octree_num_nodes(void* h) {
    auto* t = static_cast<RefTree*>(h);
    if (t->node_keys.empty()) {
        auto lt = t->tree->hashmap.lock_table();
        for (const auto& it : lt) t->node_keys.push_back(it.first);
        std::sort(t->node_keys.begin(), t->node_keys.end());
    }
    return t->node_keys.size();
}
void ref_octree_nodes(void* h, uint64_t* out) {
    auto* t = static_cast<RefTree*>(h);
    ref_octree_num_nodes(h);
    std::memcpy(out, t->node_keys.data(),
                t->node_keys.size() * sizeof(uint64_t));
}
// voxel_size[0..21], offset[3]
void ref_octree_params(void* h, float* voxel_sizes22, float* inv_voxel_sizes22,
                       int* offset3) {
    auto& tr = *static_cast<RefTree*>(h)->tree;
    for (int i = 0; i <= 21; ++i) {
        voxel_sizes22[i] = tr.VoxelSize(i);
        inv_voxel_sizes22[i] = tr.InvVoxelSize(i);
    }
    offset3[0] = tr.offset.x();
    offset3[1] = tr.offset.y();
    offset3[2] = tr.offset.z();
}

void* ref_grids_create(void* h, int num_levels, int voxel_info_all_levels) {
    try {
        auto* g = new RefGrids;
        g->grids = asr::CreateGridsFromOctree(*static_cast<RefTree*>(h)->tree,
                                              num_levels,
                                              voxel_info_all_levels != 0);
        return g;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
void ref_grids_free(void* g) { delete static_cast<RefGrids*>(g); }

// field ids: 0 voxel_keys u64, 1 voxel_centers f32, 2 voxel_sizes f32,
// 3 neighbors_index i32, 4 neighbors_kernel_index u8, 5 neighbors_row_splits
// i64, 6 up_neighbors_index i32, 7 up_neighbors_kernel_index u8,
// 8 up_neighbors_row_splits i64.  Returns element count; copies if out != 0.
uint64_t ref_grids_field(void* gh, int level, int field, void* out) {
    auto& g = *static_cast<RefGrids*>(gh)->grids.at(level);
#define FIELD(id, vec)                                                     \
    case id:                                                               \
        if (out)                                                           \
            std::memcpy(out, g.vec.data(), g.vec.size() * sizeof(g.vec[0])); \
        return g.vec.size();
    switch (field) {
        FIELD(0, voxel_keys)
        FIELD(1, voxel_centers)
        FIELD(2, voxel_sizes)
        FIELD(3, neighbors_index)
        FIELD(4, neighbors_kernel_index)
        FIELD(5, neighbors_row_splits)
        FIELD(6, up_neighbors_index)
        FIELD(7, up_neighbors_kernel_index)
        FIELD(8, up_neighbors_row_splits)
    }
#undef FIELD
    return 0;
}

void* ref_duals_create(void* h) {
    try {
        auto* d = new RefDuals;
        asr::CreateDualVertexIndices(d->idx, *static_cast<RefTree*>(h)->tree);
        return d;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
uint64_t ref_duals_size(void* d) { return static_cast<RefDuals*>(d)->idx.size(); }
void ref_duals_copy(void* d, uint64_t* out) {
    auto& v = static_cast<RefDuals*>(d)->idx;
    for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
}
void ref_duals_free(void* d) { delete static_cast<RefDuals*>(d); }

void* ref_mesh_create(const float* values, uint64_t num_voxels,
                      const uint64_t* dual_indices, uint64_t num_duals,
                      const float* node_positions, float unsigned_threshold) {
    try {
        auto* m = new RefMesh;
        std::vector<float> vals(values, values + num_voxels * 2);
        std::vector<size_t> duals(num_duals * 8);
        for (size_t i = 0; i < duals.size(); ++i) duals[i] = dual_indices[i];
        std::vector<float> pos(node_positions, node_positions + num_voxels * 3);
        asr::CreateTriangleMesh(m->vertices, m->triangles, vals, duals, pos,
                                unsigned_threshold);
        return m;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
void* ref_mesh_remove_components(const float* vertices, uint64_t nv,
                                 const int32_t* triangles, uint64_t nt,
                                 int64_t keep_n, int64_t min_size) {
    try {
        auto* m = new RefMesh;
        m->vertices.assign(vertices, vertices + nv * 3);
        m->triangles.assign(triangles, triangles + nt * 3);
        asr::RemoveConnectedComponents(m->vertices, m->triangles, keep_n,
                                       min_size);
        return m;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
uint64_t ref_mesh_num_vertices(void* m) {
    return static_cast<RefMesh*>(m)->vertices.size() / 3;
}
uint64_t ref_mesh_num_triangles(void* m) {
    return static_cast<RefMesh*>(m)->triangles.size() / 3;
}
void ref_mesh_copy(void* mh, float* vertices, int32_t* triangles) {
    auto* m = static_cast<RefMesh*>(mh);
    if (vertices)
        std::memcpy(vertices, m->vertices.data(),
                    m->vertices.size() * sizeof(float));
    if (triangles)
        std::memcpy(triangles, m->triangles.data(),
                    m->triangles.size() * sizeof(int32_t));
}
void ref_mesh_free(void* m) { delete static_cast<RefMesh*>(m); }

}  // extern "C"
