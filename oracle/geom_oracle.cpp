// TEST INFRASTRUCTURE ONLY — CPU restatement ("port" oracle) of the integer /
// geometry half of the hot path.  It is the checker for the CUDA kernels, never
// the thing measured or shipped: only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load it.
//
// Parity status: PINNED.  tests/test_oracle_geometry.py checks every output of
// this file bit-for-bit against oracle/_ref/libasr_ref.so, which is the
// reference's own cpp/lib sources compiled unmodified (oracle/Makefile).
//
// It is written independently of the reference implementation: coordinates are
// handled as integer triples with loop-based bit interleaving (not the magic
// mask Morton code), the node set is a hash set of *sibling-group* keys, and the
// slot tables are derived from geometry instead of being tabulated.  Each
// function cites the reference lines whose behaviour it restates.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace {

typedef uint64_t Key;
const int kMaxLevel = 21;  // OctreeBase::MAX_LEVEL() for 64-bit keys, octreebase.h:43

struct Cell {
    int x, y, z, lev;
};

// ---- location codes (zindex.h:34,92 ; octreebase.h:53-75) -------------------
Key interleave(uint32_t x, uint32_t y, uint32_t z) {
    Key m = 0;
    for (int b = 0; b < kMaxLevel; ++b) {
        m |= Key((x >> b) & 1u) << (3 * b);
        m |= Key((y >> b) & 1u) << (3 * b + 1);
        m |= Key((z >> b) & 1u) << (3 * b + 2);
    }
    return m;
}
int level_of(Key k) { return (63 - __builtin_clzll(k)) / 3; }  // k != 0
bool cell_ok(const Cell& c) {  // octreebase.h:120-128
    if (c.lev > kMaxLevel || c.lev < 0) return false;
    const int n = 1 << c.lev;
    return c.x >= 0 && c.x < n && c.y >= 0 && c.y < n && c.z >= 0 && c.z < n;
}
Key key_of(const Cell& c) {  // octreebase.h:59-65 ; 0 == INVALID_KEY
    if (!cell_ok(c)) return 0;
    return interleave(c.x, c.y, c.z) | (Key(1) << (3 * c.lev));
}
Cell cell_of(Key k) {  // octreebase.h:67-77
    Cell c;
    c.lev = level_of(k);
    k &= ~(Key(1) << (3 * c.lev));
    c.x = c.y = c.z = 0;
    for (int b = 0; b < kMaxLevel; ++b) {
        c.x |= int((k >> (3 * b)) & 1) << b;
        c.y |= int((k >> (3 * b + 1)) & 1) << b;
        c.z |= int((k >> (3 * b + 2)) & 1) << b;
    }
    return c;
}

// ---- octree frame (octree.cpp:20-42, octree.h:42-93) ------------------------
struct Frame {
    float vs[kMaxLevel + 1];
    float ivs[kMaxLevel + 1];
    int off[3];
    float bb_min[3], bb_max[3];
};

void init_frame(Frame& f, const float* bb_min, const float* bb_max) {
    float centre[3], edge = 0.f;
    for (int a = 0; a < 3; ++a) {
        f.bb_min[a] = bb_min[a];
        f.bb_max[a] = bb_max[a];
        centre[a] = 0.5f * (bb_max[a] + bb_min[a]);
    }
    edge = bb_max[0] - bb_min[0];
    edge = std::max(edge, bb_max[1] - bb_min[1]);
    edge = std::max(edge, bb_max[2] - bb_min[2]);
    edge *= 1.f;  // scale_bb == 1 (octree.cpp:238)
    f.vs[0] = edge;
    f.ivs[0] = 1 / edge;
    for (int i = 1; i <= kMaxLevel; ++i) {
        double t = edge * (1.0 / std::pow(2, i));
        f.vs[i] = float(t);
        f.ivs[i] = float(1.0 / t);
    }
    for (int a = 0; a < 3; ++a) {
        float lo = centre[a] - 0.5f * edge;
        f.off[a] = int(-std::floor(lo * f.ivs[kMaxLevel]));
    }
}
int level_from_scale(const Frame& f, float scale) {  // octree.h:42-47
    for (int l = 0; l <= kMaxLevel; ++l)
        if (f.vs[l] < scale) return std::max(0, l - 1);
    return kMaxLevel;
}
Cell cell_of_point(const Frame& f, const float* p, int lev) {  // octree.h:49-63
    Cell c;
    lev = std::min(kMaxLevel, lev);
    int q[3];
    for (int a = 0; a < 3; ++a) {
        q[a] = int(std::floor(p[a] * f.ivs[kMaxLevel]));
        q[a] += f.off[a];
        q[a] >>= (kMaxLevel - lev);
    }
    c.x = q[0];
    c.y = q[1];
    c.z = q[2];
    c.lev = lev;
    return c;
}
void centre_of(const Frame& f, const Cell& c, float* out) {  // octree.h:78-93
    const int s = kMaxLevel - c.lev;
    const int q[3] = {c.x << s, c.y << s, c.z << s};
    for (int a = 0; a < 3; ++a) {
        // int + double(0.5f * 2^s), times float->double voxel size, rounded once
        double t = (double(q[a] - f.off[a]) + double(0.5f) * std::pow(2, s)) *
                   double(f.vs[kMaxLevel]);
        out[a] = float(t);
    }
}

// ---- octree node set --------------------------------------------------------
// A "group" is the key of the first of 8 siblings (key & ~7).  The node set of
// the reference hash map (octree.cpp:230-280) is { 8 members of every group }
// plus the root key 1.
struct Tree {
    Frame f;
    std::unordered_set<Key> groups;
    bool any = false;
    std::vector<Key> nodes;   // sorted
    std::vector<Key> leaves;  // sorted (octree.cpp:227)
    std::unordered_map<Key, int64_t> leaf_index;  // leaf key -> position
    std::unordered_set<Key> node_set;

    bool has_node(Key k) const {
        if (k == 1) return any;
        return groups.count(k & ~Key(7)) != 0;
    }
    bool has_first_child(Key k) const {  // octreebase.h:174-182
        if (k != 0 && __builtin_clzll(k) <= 1) return false;
        return groups.count(k << 3) != 0;
    }
};

// octree.cpp:110-150 : every inserted key drags in its 7 siblings and, level by
// level, all ancestors with their siblings.
void add_with_ancestors(Tree& t, Key k) {
    t.any = true;
    while (k > 1) {
        Key g = k & ~Key(7);
        if (!t.groups.insert(g).second) return;
        k = g >> 3;
    }
    if (k == 0) t.groups.insert(0);  // INVALID_KEY inserted by octree.cpp:257
}

// octree.cpp:152-206.  Sequential sweep: the first pass visits the first-sibling keys in the
// iteration order of the reference's hash map — ascending keys for the pinned oracle (std::map
// stand-in, oracle/shim/libcuckoo) — and every later pass visits the groups created by the previous
// one in creation order (keys_to_process2, :195-197).  The leaf test (:171) sees the nodes inserted
// so far, so the order is part of the algorithm.
void balance(Tree& t) {
    static const int dirs[6][3] = {{-1, 0, 0}, {0, -1, 0}, {0, 0, -1},
                                   {1, 0, 0},  {0, 1, 0},  {0, 0, 1}};
    std::vector<Key> work(t.groups.begin(), t.groups.end()), next;
    std::sort(work.begin(), work.end());
    while (!work.empty()) {
        for (Key g : work) {
            if (g == 0) continue;  // key 0: HasFirstChild(0) -> contains(0) -> not a leaf
            if (t.has_first_child(g)) continue;
            Cell pc = cell_of(g >> 3);
            for (const auto& d : dirs) {
                Cell nc = {pc.x + d[0], pc.y + d[1], pc.z + d[2], pc.lev};
                Key k = key_of(nc);
                if (!k) continue;
                while (!t.has_node(k)) {
                    Key ng = k & ~Key(7);
                    if (t.groups.insert(ng).second) next.push_back(ng);
                    k >>= 3;
                }
            }
        }
        work.swap(next);
        next.clear();
    }
}

void finish_tree(Tree& t) {  // octree.cpp:208-228
    t.nodes.clear();
    if (t.any) t.nodes.push_back(1);
    for (Key g : t.groups)
        for (int j = 0; j < 8; ++j)
            if (g + j != 1) t.nodes.push_back(g + j);
    std::sort(t.nodes.begin(), t.nodes.end());
    t.node_set.insert(t.nodes.begin(), t.nodes.end());
    for (Key k : t.nodes)
        if (!t.has_first_child(k)) t.leaves.push_back(k);
    for (size_t i = 0; i < t.leaves.size(); ++i) t.leaf_index[t.leaves[i]] = i;
}

// ---- grids (grid.cpp:43-314) -------------------------------------------------
struct Grid {
    std::vector<Key> keys;
    std::vector<float> centres, sizes;
    std::vector<int32_t> nidx;
    std::vector<uint8_t> nslot;
    std::vector<int64_t> nsplits;
    std::vector<int32_t> uidx;
    std::vector<uint8_t> uslot;
    std::vector<int64_t> usplits;
};

int64_t find_sorted(const std::vector<Key>& keys, Key k) {
    auto it = std::lower_bound(keys.begin(), keys.end(), k);
    return (it != keys.end() && *it == k) ? int64_t(it - keys.begin()) : -1;
}

// Kernel-slot layout (grid.cpp:43-175): 0 self | 1..6 same level (-x,+x,-y,+y,
// -z,+z) | 7..30 finer: face f, 4 children, lower in-face axis fastest |
// 31..54 coarser: face f, position of the neighbour inside its parent's face.
void face_adjacency(Grid& g) {
    const auto& keys = g.keys;
    g.nidx.clear();
    g.nslot.clear();
    g.nsplits.assign(keys.size() + 1, 0);
    for (size_t i = 0; i < keys.size(); ++i) {
        const Cell c = cell_of(keys[i]);
        int count = 1;
        g.nidx.push_back(int32_t(i));
        g.nslot.push_back(0);
        for (int f = 0; f < 6; ++f) {  // same level
            int d[3] = {0, 0, 0};
            d[f / 2] = (f & 1) ? 1 : -1;
            Key nk = key_of({c.x + d[0], c.y + d[1], c.z + d[2], c.lev});
            int64_t j = nk ? find_sorted(keys, nk) : -1;
            if (j >= 0) {
                g.nidx.push_back(int32_t(j));
                g.nslot.push_back(uint8_t(1 + f));
                ++count;
            }
        }
        if (c.lev < kMaxLevel) {  // finer
            for (int f = 0; f < 6; ++f) {
                const int a = f / 2, u = (a == 0) ? 1 : 0, v = (a == 2) ? 1 : 2;
                for (int q = 0; q < 4; ++q) {
                    int p[3] = {2 * c.x, 2 * c.y, 2 * c.z};
                    p[a] += (f & 1) ? 2 : -1;
                    p[u] += q & 1;
                    p[v] += q >> 1;
                    Key nk = key_of({p[0], p[1], p[2], c.lev + 1});
                    int64_t j = nk ? find_sorted(keys, nk) : -1;
                    if (j >= 0) {
                        g.nidx.push_back(int32_t(j));
                        g.nslot.push_back(uint8_t(7 + 4 * f + q));
                        ++count;
                    }
                }
            }
        }
        if (c.lev > 0) {  // coarser
            for (int f = 0; f < 6; ++f) {
                const int a = f / 2, u = (a == 0) ? 1 : 0, v = (a == 2) ? 1 : 2;
                int p[3] = {c.x, c.y, c.z};
                p[a] += (f & 1) ? 1 : -1;
                Key nk = key_of({p[0], p[1], p[2], c.lev});
                if (!nk) continue;
                int64_t j = find_sorted(keys, nk >> 3);
                if (j < 0) continue;
                // where the same-level neighbour sits inside its parent, seen
                // from the shared face: lower in-face axis fastest.  When the
                // neighbour is not on the far side of a parent boundary the
                // reference table holds -1 (grid.cpp:58-64); that case needs the
                // cell's own parent to be in the same grid, which never happens.
                const bool crosses = ((p[a] & 1) != 0) == ((f & 1) == 0);
                int q = crosses ? ((p[u] & 1) + 2 * (p[v] & 1)) : -1;
                g.nidx.push_back(int32_t(j));
                g.nslot.push_back(uint8_t(31 + 4 * f + q));
                ++count;
            }
        }
        g.nsplits[i + 1] = g.nsplits[i] + count;
    }
}

// grid.cpp:177-243
void coarsen(const std::vector<Key>& fine, std::vector<Key>& coarse,
             std::vector<int32_t>& uidx, std::vector<uint8_t>& uslot,
             std::vector<int64_t>& usplits) {
    const size_t n = fine.size();
    std::vector<char> merged(n, 0);
    coarse.clear();
    for (size_t i = 0; i < n;) {
        bool full = (fine[i] & 7) == 0 && i + 7 < n && fine[i + 7] == fine[i] + 7;
        if (full) {
            coarse.push_back(fine[i] >> 3);
            for (int j = 0; j < 8; ++j) merged[i + j] = 1;
            i += 8;
        } else {
            coarse.push_back(fine[i]);
            ++i;
        }
    }
    std::sort(coarse.begin(), coarse.end());
    uidx.resize(n);
    uslot.resize(n);
    usplits.resize(n + 1);
    for (size_t i = 0; i <= n; ++i) usplits[i] = int64_t(i);
    for (size_t i = 0; i < n; ++i) {
        Key target = merged[i] ? (fine[i] >> 3) : fine[i];
        uidx[i] = int32_t(std::lower_bound(coarse.begin(), coarse.end(), target) -
                          coarse.begin());
        uslot[i] = merged[i] ? uint8_t(fine[i] & 7) : uint8_t(8);
    }
}

void voxel_info(const Tree& t, Grid& g) {  // grid.cpp:251-268
    g.centres.resize(3 * g.keys.size());
    g.sizes.resize(g.keys.size());
    for (size_t i = 0; i < g.keys.size(); ++i) {
        Cell c = cell_of(g.keys[i]);
        centre_of(t.f, c, &g.centres[3 * i]);
        g.sizes[i] = t.f.vs[c.lev];
    }
}

// ---- dual cells (grid.cpp:316-459) -------------------------------------------
void dual_cells(const Tree& t, std::vector<uint64_t>& out) {
    out.clear();
    for (Key leaf : t.leaves) {
        const Cell c = cell_of(leaf);
        const int n = 1 << c.lev;
        for (int i = 0; i < 8; ++i) {
            // corner i of the cell; only strictly interior octree vertices are
            // valid (octreebase.h:86-106)
            const int vx = c.x + (i & 1), vy = c.y + ((i >> 1) & 1),
                      vz = c.z + ((i >> 2) & 1);
            if (vx < 1 || vx > n - 1 || vy < 1 || vy > n - 1 || vz < 1 || vz > n - 1)
                continue;
            Key adj[8];
            for (int j = 0; j < 8; ++j)  // octreebase.h:108-118
                adj[j] = key_of({vx - (j & 1), vy - ((j >> 1) & 1),
                                 vz - ((j >> 2) & 1), c.lev});
            bool skip = false;
            for (int j = 0; j < 8 && !skip; ++j) {
                if (j == i) continue;
                if (!t.node_set.count(adj[j])) continue;           // covered by a coarser leaf
                if (!t.leaf_index.count(adj[j])) skip = true;       // a finer leaf owns the vertex
                else if (adj[j] < leaf) skip = true;                // smallest same-level leaf owns it
            }
            if (skip) continue;
            for (int j = 0; j < 8; ++j) {
                Key k = adj[j];
                while (k && !t.node_set.count(k)) k >>= 3;
                auto it = t.leaf_index.find(k);
                // reference throws here (grid.cpp:435-440); report as ~0
                out.push_back(it == t.leaf_index.end() ? ~uint64_t(0)
                                                       : uint64_t(it->second));
            }
        }
    }
}

// ---- dual contouring, vertex part (contouring.cpp:66-199) --------------------
// corner j of a dual = adjacent node j (bit0 -> -x, bit1 -> -y, bit2 -> -z);
// the 12 edges join corners that differ in exactly one bit.
const int kEdges[12][2] = {{0, 1}, {1, 3}, {3, 2}, {2, 0}, {4, 5}, {5, 7},
                           {7, 6}, {6, 4}, {0, 4}, {1, 5}, {3, 7}, {2, 6}};

bool edge_crosses(const float* values, uint64_t a, uint64_t b, float thr) {
    float u1 = values[2 * a + 1], u2 = values[2 * b + 1];
    if (u1 > thr && u2 > thr) return false;
    float s1 = values[2 * a], s2 = values[2 * b];
    return (s1 < 0 && s2 > 0) || (s1 > 0 && s2 < 0);
}

struct Mesh {
    std::vector<float> vertices;
    std::vector<uint64_t> dual_of_vertex;
};

void contour_vertices(Mesh& m, const float* values, const uint64_t* duals,
                      uint64_t num_duals, const float* pos, float thr) {
    for (uint64_t d = 0; d < num_duals; ++d) {
        const uint64_t* c = duals + 8 * d;
        double acc[3] = {0, 0, 0};
        int cnt = 0;
        for (const auto& e : kEdges) {
            uint64_t a = c[e[0]], b = c[e[1]];
            if (!edge_crosses(values, a, b, thr)) continue;
            double v1 = values[2 * a], v2 = values[2 * b];
            double tt = -v1 / (v2 - v1);
            if (!std::isfinite(tt) || tt < 0 || tt > 1) tt = 0.5;
            for (int k = 0; k < 3; ++k)
                acc[k] += (1 - tt) * double(pos[3 * a + k]) + tt * double(pos[3 * b + k]);
            ++cnt;
        }
        if (!cnt) continue;
        for (int k = 0; k < 3; ++k) m.vertices.push_back(float(acc[k] / cnt));
        m.dual_of_vertex.push_back(d);
    }
}

struct Grids {
    std::vector<Grid> g;
};

thread_local std::string g_err;

}  // namespace

extern "C" {

const char* og_last_error() { return g_err.c_str(); }

// CreateOctreeFromPoints, octree.cpp:230-280 (grow_steps must be 0 as in every
// caller: asr.cpp:153, module.cpp:373)
void* og_octree_create(const float* points, uint64_t n, const float* radii,
                       const float* bb_min, const float* bb_max,
                       float radius_scale, int grow_steps, int max_depth) {
    if (grow_steps != 0) {
        g_err = "grow_steps != 0 is not supported";
        return nullptr;
    }
    Tree* t = new Tree;
    init_frame(t->f, bb_min, bb_max);
    for (uint64_t i = 0; i < n; ++i) {
        const float* p = points + 3 * i;
        bool outside = false;
        for (int a = 0; a < 3; ++a)
            if (p[a] < bb_min[a] || p[a] > bb_max[a]) outside = true;
        if (outside) continue;
        int lev = level_from_scale(t->f, radius_scale * radii[i]);
        lev = std::min(max_depth, lev);
        Key k = key_of(cell_of_point(t->f, p, lev));
        if (k == 1) t->any = true;
        else add_with_ancestors(*t, k);
    }
    balance(*t);
    finish_tree(*t);
    return t;
}
void og_octree_free(void* h) { delete static_cast<Tree*>(h); }
uint64_t og_octree_num_leaves(void* h) { return static_cast<Tree*>(h)->leaves.size(); }
void og_octree_leaves(void* h, uint64_t* out) {
    auto& v = static_cast<Tree*>(h)->leaves;
    std::memcpy(out, v.data(), v.size() * 8);
}
uint64_t og_octree_num_nodes(void* h) { return static_cast<Tree*>(h)->nodes.size(); }
void og_octree_nodes(void* h, uint64_t* out) {
    auto& v = static_cast<Tree*>(h)->nodes;
    std::memcpy(out, v.data(), v.size() * 8);
}
void og_octree_params(void* h, float* vs, float* ivs, int* off) {
    auto& f = static_cast<Tree*>(h)->f;
    std::memcpy(vs, f.vs, sizeof(f.vs));
    std::memcpy(ivs, f.ivs, sizeof(f.ivs));
    std::memcpy(off, f.off, sizeof(f.off));
}

// CreateGridsFromOctree, grid.cpp:245-314
void* og_grids_create(void* h, int num_levels, int voxel_info_all_levels) {
    Tree& t = *static_cast<Tree*>(h);
    Grids* gs = new Grids;
    gs->g.resize(num_levels);
    gs->g[0].keys = t.leaves;
    for (int l = 0; l < num_levels; ++l) {
        Grid& g = gs->g[l];
        if (l > 0) {
            Grid& prev = gs->g[l - 1];
            coarsen(prev.keys, g.keys, prev.uidx, prev.uslot, prev.usplits);
        }
        if (l == 0 || voxel_info_all_levels) voxel_info(t, g);
        face_adjacency(g);
    }
    if (!voxel_info_all_levels)
        for (int l = 1; l < num_levels; ++l) gs->g[l].keys.clear();
    return gs;
}
void og_grids_free(void* g) { delete static_cast<Grids*>(g); }
uint64_t og_grids_field(void* gh, int level, int field, void* out) {
    Grid& g = static_cast<Grids*>(gh)->g.at(level);
#define FIELD(id, vec)                                                        \
    case id:                                                                  \
        if (out) std::memcpy(out, g.vec.data(), g.vec.size() * sizeof(g.vec[0])); \
        return g.vec.size();
    switch (field) {
        FIELD(0, keys)
        FIELD(1, centres)
        FIELD(2, sizes)
        FIELD(3, nidx)
        FIELD(4, nslot)
        FIELD(5, nsplits)
        FIELD(6, uidx)
        FIELD(7, uslot)
        FIELD(8, usplits)
    }
#undef FIELD
    return 0;
}

// CreateDualVertexIndices, grid.cpp:450-459
void* og_duals_create(void* h) {
    auto* v = new std::vector<uint64_t>;
    dual_cells(*static_cast<Tree*>(h), *v);
    for (uint64_t x : *v)
        if (x == ~uint64_t(0)) {
            g_err = "dual corner does not resolve to a leaf";
            delete v;
            return nullptr;
        }
    return v;
}
uint64_t og_duals_size(void* d) { return static_cast<std::vector<uint64_t>*>(d)->size(); }
void og_duals_copy(void* d, uint64_t* out) {
    auto& v = *static_cast<std::vector<uint64_t>*>(d);
    std::memcpy(out, v.data(), v.size() * 8);
}
void og_duals_free(void* d) { delete static_cast<std::vector<uint64_t>*>(d); }

// vertex part of CreateTriangleMesh, contouring.cpp:66-199
void* og_contour_create(const float* values, uint64_t num_voxels,
                        const uint64_t* dual_indices, uint64_t num_duals,
                        const float* node_positions, float unsigned_threshold) {
    (void)num_voxels;
    Mesh* m = new Mesh;
    contour_vertices(*m, values, dual_indices, num_duals, node_positions,
                     unsigned_threshold);
    return m;
}
uint64_t og_contour_num_vertices(void* m) {
    return static_cast<Mesh*>(m)->dual_of_vertex.size();
}
void og_contour_copy(void* mh, float* vertices, uint64_t* dual_of_vertex) {
    Mesh* m = static_cast<Mesh*>(mh);
    if (vertices) std::memcpy(vertices, m->vertices.data(), m->vertices.size() * 4);
    if (dual_of_vertex)
        std::memcpy(dual_of_vertex, m->dual_of_vertex.data(), m->dual_of_vertex.size() * 8);
}
void og_contour_free(void* m) { delete static_cast<Mesh*>(m); }

}  // extern "C"
