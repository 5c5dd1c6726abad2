// Grid hierarchy, face-adjacency tables and dual cells on the GPU.
//
// Replaces CreateGridsFromOctree / CreateLeafNeighborInformation /
// CombineSiblings (reference cpp/lib/grid.cpp:43-314) and CreateDualCells /
// CreateDualVertexIndices (:316-459).  The reference loops over voxels on one
// thread with a std::lower_bound per probe; here one warp owns a voxel, its
// lanes run the 37 independent probes (1 self + 6 same-level + 24 finer + 6
// coarser) in parallel, a warp OR-reduction yields the 55-bit slot mask, and a
// scan over popcounts gives the CSR offsets (count -> scan -> fill).
#include "hash.cuh"
#include "internal.h"
#include "prims.cuh"
#include "profile.cuh"

namespace asrb {

// ------------------------------------------------------------------ face adjacency
// Kernel-slot layout (grid.cpp:43-175): 0 self | 1..6 same level (-x,+x,-y,+y,-z,+z)
// | 7..30 finer: face f, four children, lower in-face axis fastest | 31..54 coarser:
// face f, position of this voxel's same-level neighbour inside its parent.
struct Probe {
    Key key;   // key to look up (0 = nothing to look up)
    int slot;  // kernel slot if found
};

__device__ __forceinline__ Probe make_probe(const Cell& c, int p) {
    Probe r{0, 0};
    if (p == 0) {
        r.key = cell_key(c.x, c.y, c.z, c.lev);
        r.slot = 0;
    } else if (p < 7) {
        const int f = p - 1;
        int d[3] = {0, 0, 0};
        d[f >> 1] = (f & 1) ? 1 : -1;
        r.key = cell_key(c.x + d[0], c.y + d[1], c.z + d[2], c.lev);
        r.slot = p;
    } else if (p < 31) {
        if (c.lev >= kMaxLevel) return r;
        const int j = p - 7, f = j >> 2, q = j & 3;
        const int a = f >> 1, u = (a == 0) ? 1 : 0, v = (a == 2) ? 1 : 2;
        int pos[3] = {2 * c.x, 2 * c.y, 2 * c.z};
        pos[a] += (f & 1) ? 2 : -1;
        pos[u] += q & 1;
        pos[v] += q >> 1;
        r.key = cell_key(pos[0], pos[1], pos[2], c.lev + 1);
        r.slot = p;
    } else {
        if (c.lev <= 0) return r;
        const int f = p - 31;
        const int a = f >> 1, u = (a == 0) ? 1 : 0, v = (a == 2) ? 1 : 2;
        int pos[3] = {c.x, c.y, c.z};
        pos[a] += (f & 1) ? 1 : -1;
        const Key nk = cell_key(pos[0], pos[1], pos[2], c.lev);
        if (!nk) return r;
        // the reference's table value (-1 when the neighbour shares our parent,
        // grid.cpp:58-64) is reproduced arithmetically; that case never finds a
        // node because a voxel and its parent are never in the same grid.
        const bool crosses = ((pos[a] & 1) != 0) == ((f & 1) == 0);
        const int q = crosses ? ((pos[u] & 1) + 2 * (pos[v] & 1)) : -1;
        r.key = nk >> 3;
        r.slot = 31 + 4 * f + q;
    }
    return r;
}

// A grid level is a set of non-overlapping voxels (octree leaves, then complete sibling
// groups merged into their parent), so across one face a voxel sees EITHER a same-level
// neighbour OR up to four finer ones OR one coarser one.  The probes therefore run in two
// rounds: lanes 1..6 look up the same-level neighbours, and only the faces that came back
// empty (~20 %) issue their 4 finer + 1 coarser probes (lanes 0..29 of the second round).
// The reference probes all 55 slots unconditionally (grid.cpp:102-170); the result is the
// same table (bit-exact parity tests, incl. the margin-free boxes with their junk leaves).
__device__ __forceinline__ int second_round_face(int lane) {  // lane 0..29 -> probe 7..36 -> face
    const int p = lane + 7;
    return p < 31 ? (p - 7) >> 2 : p - 31;
}

// four lookups of a lane with all first probes in flight together (key 0 = no lookup)
__device__ __forceinline__ void table_find4(const KeyTableView t, const Key (&k)[4], long long (&out)[4]) {
    uint32_t s[4];
    ulonglong2 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        s[j] = hash_key(k[j]) & t.mask;
        v[j] = make_ulonglong2(kNoKey, 0);
        if (k[j]) v[j] = __ldg(reinterpret_cast<const ulonglong2*>(t.e + s[j]));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        out[j] = -1;
        if (!k[j]) continue;
        for (;;) {
            if (v[j].x == k[j]) {
                out[j] = (long long)v[j].y;
                break;
            }
            if (v[j].x == kNoKey) break;
            s[j] = (s[j] + 1) & t.mask;
            v[j] = __ldg(reinterpret_cast<const ulonglong2*>(t.e + s[j]));
        }
    }
}

// pass 1: 55-bit slot mask per voxel; the indices found are kept, in slot order, in `stash`
// ([V][32] int32: a row has at most 1 + 6 x 4 = 25 entries), so pass 2 is a plain copy.
// A warp owns FOUR consecutive voxels (the kernel is bound by the latency of dependent hash probes: round 2's
// version with one voxel per warp had 6 probes in flight per warp in round 1): round 1 = the 4 x 6 same-level
// probes at once (octet o of the warp = voxel o, lanes 1..6 of it = faces), round 2 = up to 4 x 30 probes, four
// per lane issued together.
__global__ void __launch_bounds__(256)
adjacency_mask_kernel(const Key* __restrict__ keys, long long V, const KeyTableView table,
                      unsigned long long* __restrict__ mask, int32_t* __restrict__ count, int32_t* __restrict__ stash) {
    const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long v0 = w * 4;
    if (v0 >= V) return;
    const int o = lane >> 3, f = lane & 7;
    long long i1 = -1;  // what this lane found in round 1 (lane 8 o: the voxel itself)
    if (v0 + o < V) {
        if (f == 0) {
            i1 = v0 + o;
        } else if (f < 7) {
            const Probe pr = make_probe(key_cell(keys[v0 + o]), f);
            if (pr.key) i1 = table_find(table, pr.key);
        }
    }
    const unsigned found1 = __ballot_sync(0xffffffffu, i1 >= 0);  // bit 8 o + f
    Key k2[4];
    int s2[4];
    long long i2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        k2[j] = 0;
        s2[j] = 0;
        if (v0 + j < V && lane < 30) {
            const unsigned same_faces = (found1 >> (8 * j + 1)) & 0x3fu;  // bit f - 1: face f has a same-level neighbour
            if (!((same_faces >> second_round_face(lane)) & 1)) {
                const Probe pr = make_probe(key_cell(keys[v0 + j]), lane + 7);
                k2[j] = pr.key;
                s2[j] = pr.slot;
            }
        }
    }
    table_find4(table, k2, i2);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (v0 + j >= V) break;  // warp-uniform
        unsigned long long m = 0;
        const bool mine1 = o == j && i1 >= 0;
        if (mine1) m |= 1ULL << f;
        if (i2[j] >= 0) m |= 1ULL << s2[j];
        const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)m);
        const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(m >> 32));
        m = ((unsigned long long)hi << 32) | lo;
        int32_t* row = stash + (v0 + j) * 32;
        if (mine1) row[__popcll(m & ((1ULL << f) - 1))] = (int32_t)i1;
        if (i2[j] >= 0) row[__popcll(m & ((1ULL << s2[j]) - 1))] = (int32_t)i2[j];
        if (lane == 0) {
            mask[v0 + j] = m;
            count[v0 + j] = __popcll(m);
        }
    }
}

// pass 2: (index, slot) in slot order from the stash and the mask
__global__ void __launch_bounds__(256)
adjacency_fill_kernel(long long V, const unsigned long long* __restrict__ mask, const int32_t* __restrict__ stash,
                      const int64_t* __restrict__ splits, int32_t* __restrict__ nidx, uint8_t* __restrict__ nslot) {
    const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= V) return;
    const unsigned long long m = mask[w];
    const int cnt = __popcll(m);
    if (lane >= cnt) return;
    const unsigned lo = (unsigned)m, hi = (unsigned)(m >> 32);
    const int nlo = __popc(lo);
    const int slot = lane < nlo ? (int)__fns(lo, 0, lane + 1) : 32 + (int)__fns(hi, 0, lane - nlo + 1);
    const int64_t base = splits[w];
    nidx[base + lane] = stash[w * 32 + lane];
    nslot[base + lane] = (uint8_t)slot;
}

// ------------------------------------------------------------------ coarsening
__device__ __forceinline__ bool is_merged(const Key* __restrict__ keys, long long V, long long i) {
    const Key k = keys[i];
    const long long b = i - (long long)(k & 7);
    if (b < 0 || b + 7 >= V) return false;
    return keys[b] == (k & ~Key(7)) && keys[b + 7] == (k | Key(7));
}

// fine voxel -> coarse key it maps to (kNoKey for siblings 1..7 of a merged group)
__global__ void __launch_bounds__(256)
coarsen_keys_kernel(const Key* __restrict__ keys, long long V, Key* __restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= V) return;
    const Key k = keys[i];
    Key o = k;
    if (is_merged(keys, V, i)) o = (k & 7) == 0 ? (k >> 3) : kNoKey;
    out[i] = o;
}

__global__ void __launch_bounds__(256)
up_table_kernel(const Key* __restrict__ keys, long long V, const KeyTableView coarse,
                int32_t* __restrict__ uidx, uint8_t* __restrict__ uslot) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= V) return;
    const Key k = keys[i];
    const bool m = is_merged(keys, V, i);
    uidx[i] = (int32_t)table_find(coarse, m ? (k >> 3) : k);  // always present (grid.cpp:211-215)
    uslot[i] = m ? (uint8_t)(k & 7) : (uint8_t)8;
}

// ------------------------------------------------------------------ voxel geometry
// centre = ((coord << s) - offset + 0.5 * 2^s) * voxel_size[21], evaluated in
// double and rounded once (octree.h:78-93); size = voxel_size[level].
struct VoxelFrame {
    float vs[kMaxLevel + 1];
    int off[3];
};
__global__ void __launch_bounds__(256)
voxel_info_kernel(const Key* __restrict__ keys, long long V, VoxelFrame f, float* __restrict__ centers,
                  float* __restrict__ sizes) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= V) return;
    const Cell c = key_cell(keys[i]);
    const int s = kMaxLevel - c.lev;
    const double half = 0.5 * (double)(1 << s);
    const double h = (double)f.vs[kMaxLevel];
    centers[3 * i + 0] = (float)(((double)((c.x << s) - f.off[0]) + half) * h);
    centers[3 * i + 1] = (float)(((double)((c.y << s) - f.off[1]) + half) * h);
    centers[3 * i + 2] = (float)(((double)((c.z << s) - f.off[2]) + half) * h);
    sizes[i] = f.vs[c.lev];
}

// position of the first kNoKey in a sorted array = number of surviving keys
__global__ void first_nokey_kernel(const Key* __restrict__ a, long long n, int64_t* out) {
    *out = lower_bound_key(a, n, kNoKey);
}

static void build_adjacency(GridLevel& g, const KeyTable& table, cudaStream_t s) {
    const long long V = g.V;
    g.nsplits.alloc((size_t)V + 1, s);
    DevBuf<unsigned long long> mask((size_t)V, s);
    DevBuf<int32_t> count((size_t)V, s);
    DevBuf<int32_t> stash((size_t)V * 32, s);
    if (V) {
        ProfileScope prof("adjacency_mask", s);
        adjacency_mask_kernel<<<grid_for((size_t)((V + 3) / 4) * 32, 256), 256, 0, s>>>(g.keys.get(), V, table.view(), mask.get(),
                                                                            count.get(), stash.get());
        ASRB_CHECK_LAUNCH();
    }
    exclusive_sum_i32_to_i64(count.get(), g.nsplits.get(), (size_t)V, s);
    g.E = d2h_scalar(g.nsplits.get() + V, s);
    g.nidx.alloc((size_t)g.E, s);
    g.nslot.alloc((size_t)g.E, s);
    if (V) {
        ProfileScope prof("adjacency_fill", s);
        adjacency_fill_kernel<<<grid_for((size_t)V * 32, 256), 256, 0, s>>>(V, mask.get(), stash.get(), g.nsplits.get(),
                                                                            g.nidx.get(), g.nslot.get());
        ASRB_CHECK_LAUNCH();
    }
}

void grids_build(Octree& t, int num_levels, bool all_info, cudaStream_t s) {
    ASRB_REQUIRE(num_levels >= 1, "num_levels must be >= 1");
    t.grids.clear();
    t.grids_all_info = all_info;
    VoxelFrame vf;
    for (int l = 0; l <= kMaxLevel; ++l) vf.vs[l] = t.frame.vs[l];
    for (int a = 0; a < 3; ++a) vf.off[a] = t.frame.off[a];

    KeyTable table;  // keys of the level being built -> position
    for (int l = 0; l < num_levels; ++l) {
        auto g = std::make_unique<GridLevel>();
        if (l == 0) {
            g->V = t.num_leaves;
            g->keys.alloc((size_t)g->V, s);
            if (g->V)
                ASRB_CUDA(cudaMemcpyAsync(g->keys.get(), t.leaves.get(), (size_t)g->V * sizeof(Key),
                                          cudaMemcpyDeviceToDevice, s));
            table.build(g->keys.get(), (size_t)g->V, s);
        } else {
            GridLevel& prev = *t.grids.back();
            const long long V = prev.V;
            DevBuf<Key> ck((size_t)V, s);
            if (V) {
                coarsen_keys_kernel<<<grid_for(V, 256), 256, 0, s>>>(prev.keys.get(), V, ck.get());
                ASRB_CHECK_LAUNCH();
            }
            sort_keys_u64(ck.get(), (size_t)V, s);  // kNoKey entries sort last
            // number of surviving keys = V - 7 * merged groups; count via lower bound of kNoKey
            size_t Vc = (size_t)V;
            if (V) {
                DevBuf<int64_t> pos(1, s);
                first_nokey_kernel<<<1, 1, 0, s>>>(ck.get(), V, pos.get());
                ASRB_CHECK_LAUNCH();
                Vc = (size_t)d2h_scalar(pos.get(), s);
            }
            g->V = (int64_t)Vc;
            g->keys.alloc(Vc, s);
            if (Vc) ASRB_CUDA(cudaMemcpyAsync(g->keys.get(), ck.get(), Vc * sizeof(Key), cudaMemcpyDeviceToDevice, s));
            prev.uidx.alloc((size_t)V, s);
            prev.uslot.alloc((size_t)V, s);
            prev.has_up = true;
            table.build(g->keys.get(), Vc, s);
            if (V) {
                up_table_kernel<<<grid_for(V, 256), 256, 0, s>>>(prev.keys.get(), V, table.view(), prev.uidx.get(),
                                                                 prev.uslot.get());
                ASRB_CHECK_LAUNCH();
            }
        }
        if (l == 0 || all_info) {
            g->centers.alloc((size_t)g->V * 3, s);
            g->sizes.alloc((size_t)g->V, s);
            if (g->V) {
                voxel_info_kernel<<<grid_for(g->V, 256), 256, 0, s>>>(g->keys.get(), g->V, vf, g->centers.get(),
                                                                      g->sizes.get());
                ASRB_CHECK_LAUNCH();
            }
        }
        build_adjacency(*g, table, s);
        t.grids.push_back(std::move(g));
    }
}

// ------------------------------------------------------------------ dual cells
// node lookup against the sorted sibling groups: returns node index or -1
__device__ __forceinline__ long long node_index(const KeyTableView groups, int root_separate, Key k) {
    if (k == 1 && root_separate) return 0;
    const long long gi = table_find(groups, k & ~Key(7));
    if (gi < 0) return -1;
    return gi * 8 + (long long)(k & 7) + (root_separate ? 1 : 0);
}

// vertex i of a cell is valid only strictly inside the cube (octreebase.h:86-106)
__device__ __forceinline__ bool dual_corner(const Cell& c, int i, int& vx, int& vy, int& vz) {
    const int n = 1 << c.lev;
    vx = c.x + (i & 1);
    vy = c.y + ((i >> 1) & 1);
    vz = c.z + ((i >> 2) & 1);
    return vx >= 1 && vx <= n - 1 && vy >= 1 && vy <= n - 1 && vz >= 1 && vz <= n - 1;
}

// thread = (leaf, corner); 8 consecutive lanes share a leaf.  A corner emits a
// dual cell unless one of the 7 other same-level cells around it is an interior
// node (a finer leaf owns the vertex) or a leaf with a smaller key (grid.cpp:334-360).
__global__ void __launch_bounds__(256)
dual_flag_kernel(const Key* __restrict__ leaves, long long V, const KeyTableView groups, int root_separate, const uint8_t* __restrict__ node_leaf, uint8_t* __restrict__ mask,
                 uint8_t* __restrict__ count) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long leaf_i = t >> 3;
    const int i = (int)(t & 7);
    bool emit = false;
    if (leaf_i < V) {
        const Key leaf = leaves[leaf_i];
        const Cell c = key_cell(leaf);
        int vx, vy, vz;
        if (leaf != 0 && dual_corner(c, i, vx, vy, vz)) {
            emit = true;
            for (int j = 0; j < 8 && emit; ++j) {
                if (j == i) continue;
                const Key a = cell_key(vx - (j & 1), vy - ((j >> 1) & 1), vz - ((j >> 2) & 1), c.lev);
                const long long ni = node_index(groups, root_separate, a);
                if (ni < 0) continue;
                if (!node_leaf[ni] || a < leaf) emit = false;
            }
        }
    }
    const unsigned b = __ballot_sync(0xffffffffu, emit);
    if (leaf_i < V && i == 0) {
        const unsigned m = (b >> ((threadIdx.x & 31) & ~7)) & 0xffu;
        mask[leaf_i] = (uint8_t)m;
        count[leaf_i] = (uint8_t)__popc(m);
    }
}

__global__ void __launch_bounds__(256)
dual_fill_kernel(const Key* __restrict__ leaves, long long V, const KeyTableView groups, int root_separate, const uint8_t* __restrict__ node_leaf, const int64_t* __restrict__ node_rank,
                 const uint8_t* __restrict__ mask, const int64_t* __restrict__ offset, int64_t* __restrict__ out,
                 int* __restrict__ error) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long leaf_i = t >> 3;
    const int i = (int)(t & 7);
    if (leaf_i >= V) return;
    const unsigned m = mask[leaf_i];
    if (!((m >> i) & 1)) return;
    const Cell c = key_cell(leaves[leaf_i]);
    int vx, vy, vz;
    dual_corner(c, i, vx, vy, vz);
    int64_t* row = out + (offset[leaf_i] + __popc(m & ((1u << i) - 1))) * 8;
    for (int j = 0; j < 8; ++j) {
        Key a = cell_key(vx - (j & 1), vy - ((j >> 1) & 1), vz - ((j >> 2) & 1), c.lev);
        long long ni = -1;
        while (a && (ni = node_index(groups, root_separate, a)) < 0) a >>= 3;  // walk up (grid.cpp:429-433)
        if (ni < 0 || !node_leaf[ni]) {
            *error = 1;  // the reference throws here (grid.cpp:434-440)
            row[j] = 0;
        } else {
            row[j] = node_rank[ni];
        }
    }
}

Octree::~Octree() {
    if (dual_count_event) {
        cudaEventSynchronize(dual_count_event);
        cudaEventDestroy(dual_count_event);
    }
    if (dual_err_event) {
        cudaEventSynchronize(dual_err_event);
        cudaEventDestroy(dual_err_event);
    }
    pinned_slot_release(dual_count_host);
    pinned_slot_release(dual_err_host);
}

void duals_begin(Octree& t, cudaStream_t s) {
    if (t.num_duals >= 0 || t.dual_begun) return;
    const long long V = t.num_leaves;
    t.dual_mask.alloc((size_t)V, s);
    t.dual_offset.alloc((size_t)V + 1, s);
    DevBuf<uint8_t> count((size_t)V, s);
    if (V) {
        ProfileScope prof("dual_flag", s);
        dual_flag_kernel<<<grid_for((size_t)V * 8, 256), 256, 0, s>>>(t.leaves.get(), V, t.group_table.view(),
                                                                      t.root_separate, t.node_leaf.get(),
                                                                      t.dual_mask.get(), count.get());
        ASRB_CHECK_LAUNCH();
    }
    exclusive_sum_u8_to_i64(count.get(), t.dual_offset.get(), (size_t)V, s);
    if (!t.dual_count_host) t.dual_count_host = pinned_slot_acquire();
    if (!t.dual_count_event) ASRB_CUDA(cudaEventCreateWithFlags(&t.dual_count_event, cudaEventDisableTiming));
    ASRB_CUDA(cudaMemcpyAsync(t.dual_count_host, t.dual_offset.get() + V, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    ASRB_CUDA(cudaEventRecord(t.dual_count_event, s));
    t.dual_begun = true;
}

void duals_count(Octree& t, cudaStream_t s) {
    if (t.num_duals >= 0) return;
    duals_begin(t, s);
    ASRB_CUDA(cudaEventSynchronize(t.dual_count_event));
    t.num_duals = *t.dual_count_host;
}

// the fill is queued without a host synchronisation; its error flag (the reference's "found node is not a leaf",
// grid.cpp:436-440) is read by duals_check
void duals_fill(Octree& t, int64_t* d_out, cudaStream_t s) {
    duals_count(t, s);
    const long long V = t.num_leaves;
    if (!V || !t.num_duals) return;
    t.dual_err.alloc(1, s);
    ASRB_CUDA(cudaMemsetAsync(t.dual_err.get(), 0, sizeof(int), s));
    {
        ProfileScope prof("dual_fill", s);
        dual_fill_kernel<<<grid_for((size_t)V * 8, 256), 256, 0, s>>>(t.leaves.get(), V, t.group_table.view(),
                                                                      t.root_separate, t.node_leaf.get(),
                                                                      t.node_rank.get(), t.dual_mask.get(),
                                                                      t.dual_offset.get(), d_out, t.dual_err.get());
        ASRB_CHECK_LAUNCH();
    }
    if (!t.dual_err_host) t.dual_err_host = pinned_slot_acquire();
    if (!t.dual_err_event) ASRB_CUDA(cudaEventCreateWithFlags(&t.dual_err_event, cudaEventDisableTiming));
    *t.dual_err_host = 0;
    ASRB_CUDA(cudaMemcpyAsync(t.dual_err_host, t.dual_err.get(), sizeof(int), cudaMemcpyDeviceToHost, s));
    ASRB_CUDA(cudaEventRecord(t.dual_err_event, s));
    t.dual_err_pending = true;
}

void duals_check(Octree& t) {
    if (!t.dual_err_pending) return;
    ASRB_CUDA(cudaEventSynchronize(t.dual_err_event));
    t.dual_err_pending = false;
    if ((int)(*t.dual_err_host & 0xffffffff)) throw Error(kRuntimeError, "found node is not a leaf");
}

}  // namespace asrb
