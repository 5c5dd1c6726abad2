// Adaptive octree construction on the GPU.
//
// Replaces CreateOctreeFromPoints (reference cpp/lib/octree.cpp:230-280) and the
// three passes it runs (CreateAncestorsAndSiblings :110-150, BalanceFaces
// :152-206, InitAttributesAndLeaves :208-228).  The reference walks a
// concurrent hash map node by node on one thread; here the node set is a sorted
// array of *sibling-group* keys (key & ~7, one entry per 8 nodes) and every pass
// is a data-parallel map + radix sort + unique:
//
//   points --point_group_kernel--> group key per point --sort/unique--> base groups
//   base groups --ancestor_kernel--> <=20 ancestor groups each --sort/unique--> closure
//   repeat: leaf first-siblings probe the 6 face neighbours of their parent (hash probes of a
//           snapshot), missing chains are appended atomically with their position in the reference's
//           processing order, sorted, uniqued and merged; the rare keys whose leaf test the
//           reference's SEQUENTIAL sweep would have decided differently (children created earlier in
//           the same pass) are found and resolved afterwards  (2:1 balance, see octree_build)
//   nodes = root + 8 x groups; leaf flag = first child group absent; scan -> leaves
//
// Results are bit-identical to the reference (tests/test_geometry_parity.py).
#include "internal.h"
#include "prims.cuh"
#include "profile.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <ctime>
#include <vector>

namespace asrb {

// -------------------------------------------------------------------------------------
// Frame (host).  Operation order follows octree.cpp:20-42 exactly: float centre and
// edge, double intermediate for the per-level sizes, float product before floor.
Frame make_frame(const float* bb_min, const float* bb_max) {
    Frame f;
    float edge = bb_max[0] - bb_min[0];
    edge = std::max(edge, bb_max[1] - bb_min[1]);
    edge = std::max(edge, bb_max[2] - bb_min[2]);
    f.vs[0] = edge;
    f.ivs[0] = 1 / edge;
    for (int l = 1; l <= kMaxLevel; ++l) {
        const double size = edge * (1.0 / std::pow(2, l));
        f.vs[l] = (float)size;
        f.ivs[l] = (float)(1.0 / size);
    }
    for (int a = 0; a < 3; ++a) {
        f.bb_min[a] = bb_min[a];
        f.bb_max[a] = bb_max[a];
        const float centre = 0.5f * (bb_max[a] + bb_min[a]);
        const float lo = centre - 0.5f * edge;
        f.off[a] = (int)(-std::floor(lo * f.ivs[kMaxLevel]));
    }
    return f;
}

// Key used for points that do not contribute a node (outside the box, or the
// root itself); sorts last and is stripped after the unique pass.
constexpr Key kSkip = ~Key(0);

struct FrameDev {
    float vs[kMaxLevel + 1];
    float inv_finest;
    int off[3];
    float bb_min[3], bb_max[3];
};

// One thread per point: level from the scaled radius (octree.h:42-47), integer
// cell from a float multiply + floor (octree.h:49-63), location code, and the
// sibling-group key.  __fmul_rn keeps the products un-fused like the host code.
__global__ void __launch_bounds__(256)
point_group_kernel(const float* __restrict__ points, const float* __restrict__ radii, long long n,
                   FrameDev f, float radius_scale, int max_depth, Key* __restrict__ out_group,
                   int* __restrict__ any_flag) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float px = points[3 * i], py = points[3 * i + 1], pz = points[3 * i + 2];
    Key g = kSkip;
    const bool outside = px < f.bb_min[0] || py < f.bb_min[1] || pz < f.bb_min[2] ||
                         px > f.bb_max[0] || py > f.bb_max[1] || pz > f.bb_max[2];
    if (!outside) {
        const float scale = __fmul_rn(radius_scale, radii[i]);
        int lev = kMaxLevel;
        for (int l = 0; l <= kMaxLevel; ++l)
            if (f.vs[l] < scale) {
                lev = max(0, l - 1);
                break;
            }
        lev = min(max_depth, lev);
        lev = min(lev, kMaxLevel);
        const int sh = kMaxLevel - lev;
        const int x = ((int)floorf(__fmul_rn(px, f.inv_finest)) + f.off[0]) >> sh;
        const int y = ((int)floorf(__fmul_rn(py, f.inv_finest)) + f.off[1]) >> sh;
        const int z = ((int)floorf(__fmul_rn(pz, f.inv_finest)) + f.off[2]) >> sh;
        const Key k = cell_key(x, y, z, lev);
        // k == 1: the root, no sibling group.  k == 0: INVALID_KEY, which the
        // reference inserts as a node (octree.cpp:257) -> group 0.
        if (k != 1) g = k & ~Key(7);
        *any_flag = 1;
    }
    out_group[i] = g;
}

// Every group drags in the groups of all its ancestors (octree.cpp:118-148).
// Row i of `out` receives up to 20 ancestor groups of groups[i], kSkip padded.
__global__ void __launch_bounds__(256)
ancestor_kernel(const Key* __restrict__ groups, long long n, Key* __restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Key k = groups[i] >> 3;  // parent node of the group
    Key* row = out + i * (kMaxLevel - 1);
    for (int j = 0; j < kMaxLevel - 1; ++j) {
        Key g = kSkip;
        if (k > 1) {
            g = k & ~Key(7);
            k = g >> 3;
        }
        row[j] = g;
    }
}

__device__ __forceinline__ bool has_group(const KeyTableView groups, Key g) { return table_find(groups, g) >= 0; }

// One round of face balancing (octree.cpp:152-206), evaluated against a snapshot
// of the node set.  Thread = (frontier group, face).  A group whose first
// sibling is a leaf requires the 6 face neighbours of its parent to exist; for a
// missing neighbour the chain of sibling groups up to the first existing
// ancestor is appended to `out`.
// rank of an insertion event in the reference's sequential sweep: (position of the processed key in
// the pass, face, step of the chain walk) — the order in which octree.cpp:178-196 would create nodes
constexpr unsigned long long kRankPerKey = 6ULL * 32ULL;

__device__ __forceinline__ bool first_sibling_is_leaf(const KeyTableView groups, Key g) {
    if (g == 0) return false;  // key 0 counts as having a first child (itself)
    // leaf test on the first sibling (octreebase.h:174-182)
    return !(__clzll((long long)g) > 1 && has_group(groups, g << 3));
}

__global__ void __launch_bounds__(256)
balance_round_kernel(const Key* __restrict__ frontier, long long nf, const KeyTableView groups,
                     Key* __restrict__ out, unsigned long long* __restrict__ out_rank, unsigned long long cap,
                     unsigned long long* __restrict__ out_count) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= nf * 6) return;
    const Key g = frontier[t / 6];
    const int face = (int)(t % 6);
    if (!first_sibling_is_leaf(groups, g)) return;
    const Cell pc = key_cell(g >> 3);
    int d[3] = {0, 0, 0};
    d[face % 3] = face < 3 ? -1 : 1;
    Key k = cell_key(pc.x + d[0], pc.y + d[1], pc.z + d[2], pc.lev);
    if (!k) return;
    unsigned long long step = 0;
    while (k > 1) {
        const Key kg = k & ~Key(7);
        if (has_group(groups, kg)) break;
        const unsigned long long pos = atomicAdd(out_count, 1ULL);
        if (pos < cap) {
            out[pos] = kg;
            out_rank[pos] = (unsigned long long)t * 32ULL + step;
        }
        ++step;
        k >>= 3;
    }
}

// first record of every run of equal keys (records sorted by (key, rank)): the group and its earliest creation
__global__ void __launch_bounds__(256)
record_heads_kernel(const Key* __restrict__ key, const unsigned long long* __restrict__ rank, long long n,
                    Key* __restrict__ out_key, unsigned long long* __restrict__ out_rank,
                    unsigned long long* __restrict__ count) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i > 0 && key[i - 1] == key[i]) return;
    const unsigned long long pos = atomicAdd(count, 1ULL);
    out_key[pos] = key[i];
    out_rank[pos] = rank[i];
}

// keys of the pass that are leaves in the snapshot but whose children are created during this very pass: the
// sequential sweep skips such a key if the creation comes first (octree.cpp:171) -> candidates for the host
__global__ void __launch_bounds__(256)
balance_candidates_kernel(const Key* __restrict__ frontier, long long nf, const KeyTableView groups,
                          const Key* __restrict__ new_keys, long long nn, long long* __restrict__ cand_pos,
                          Key* __restrict__ cand_key, unsigned long long cap, unsigned long long* __restrict__ count) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= nf) return;
    const Key g = frontier[i];
    if (__clzll((long long)g) <= 1 || !first_sibling_is_leaf(groups, g)) return;
    if (find_key(new_keys, nn, g << 3) < 0) return;
    const unsigned long long pos = atomicAdd(count, 1ULL);
    if (pos < cap) {
        cand_pos[pos] = i;
        cand_key[pos] = g;
    }
}

__global__ void __launch_bounds__(256)
leaf_flag_kernel(const Key* __restrict__ groups, const KeyTableView table, int root_separate, long long num_nodes,
                 uint8_t* __restrict__ flag) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= num_nodes) return;
    Key k;
    if (root_separate) k = i == 0 ? Key(1) : groups[(i - 1) >> 3] + Key((i - 1) & 7);
    else k = groups[i >> 3] + Key(i & 7);
    bool child = false;
    if (k == 0) child = true;  // contains(0 << 3)
    else if (__clzll((long long)k) > 1) child = has_group(table, k << 3);
    flag[i] = child ? 0 : 1;
}

__global__ void __launch_bounds__(256)
leaf_compact_kernel(const Key* __restrict__ groups, int root_separate, long long num_nodes,
                    const uint8_t* __restrict__ flag, const int64_t* __restrict__ rank, Key* __restrict__ leaves) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= num_nodes || !flag[i]) return;
    Key k;
    if (root_separate) k = i == 0 ? Key(1) : groups[(i - 1) >> 3] + Key((i - 1) & 7);
    else k = groups[i >> 3] + Key(i & 7);
    leaves[rank[i]] = k;
}

// strips trailing kSkip entries of a sorted, uniqued array
static size_t strip_skip(const Key* d, size_t n, cudaStream_t s) {
    if (n == 0) return 0;
    Key last = d2h_scalar(d + n - 1, s);
    return last == kSkip ? n - 1 : n;
}

void octree_build(Octree& t, const float* d_points, const float* d_radii, int64_t n, float radius_scale,
                  int max_depth, cudaStream_t s) {
    PhaseTimer pt(s);
    FrameDev fd;
    for (int l = 0; l <= kMaxLevel; ++l) fd.vs[l] = t.frame.vs[l];
    fd.inv_finest = t.frame.ivs[kMaxLevel];
    for (int a = 0; a < 3; ++a) {
        fd.off[a] = t.frame.off[a];
        fd.bb_min[a] = t.frame.bb_min[a];
        fd.bb_max[a] = t.frame.bb_max[a];
    }

    // 1. per-point group keys -> base groups
    DevBuf<Key> pk((size_t)std::max<int64_t>(n, 1), s);
    DevBuf<int> any(1, s);
    ASRB_CUDA(cudaMemsetAsync(any.get(), 0, sizeof(int), s));
    if (n > 0) {
        point_group_kernel<<<grid_for(n, 256), 256, 0, s>>>(d_points, d_radii, n, fd, radius_scale, max_depth,
                                                            pk.get(), any.get());
        ASRB_CHECK_LAUNCH();
    }
    sort_keys_u64(pk.get(), (size_t)n, s);
    size_t nb = unique_u64(pk.get(), (size_t)n, s);
    nb = strip_skip(pk.get(), nb, s);
    t.any = d2h_scalar(any.get(), s) != 0;
    pt.lap("octree: point groups", (long long)nb);

    // 2. closure under "parent exists, with all its siblings"
    const int A = kMaxLevel - 1;
    DevBuf<Key> all(nb * (A + 1) + 1, s);
    if (nb) {
        ASRB_CUDA(cudaMemcpyAsync(all.get(), pk.get(), nb * sizeof(Key), cudaMemcpyDeviceToDevice, s));
        ancestor_kernel<<<grid_for(nb, 256), 256, 0, s>>>(pk.get(), (long long)nb, all.get() + nb);
        ASRB_CHECK_LAUNCH();
    }
    pk.release();
    size_t ng = nb * (A + 1);
    sort_keys_u64(all.get(), ng, s);
    ng = unique_u64(all.get(), ng, s);
    ng = strip_skip(all.get(), ng, s);

    DevBuf<Key> groups(ng, s);
    if (ng) ASRB_CUDA(cudaMemcpyAsync(groups.get(), all.get(), ng * sizeof(Key), cudaMemcpyDeviceToDevice, s));
    all.release();
    pt.lap("octree: ancestor closure", (long long)ng);

    // 3. 2:1 face balance to a fixed point (octree.cpp:152-206).
    // The reference sweeps the first-sibling keys SEQUENTIALLY — pass 1 in the iteration order of its hash map
    // (ascending keys for the pinned oracle, oracle/shim/libcuckoo), every later pass in the order in which the
    // previous pass created groups — and tests "is this key a leaf?" against the nodes inserted so far.  A pass
    // here evaluates all keys against a snapshot and records, for every missing group, the rank (key position,
    // face, chain step) of each event that would create it; the earliest rank per group gives the creation order
    // (= next pass's order).  A key whose children are created by an EARLIER event of the same pass would have been
    // skipped by the sequential sweep: such keys are rare (17 at 10 M points, none below a few million), they are
    // detected on the device and resolved on the host in sweep order; their events are then withdrawn.
    DevBuf<Key> frontier(ng, s);
    size_t nf = ng;
    if (ng) ASRB_CUDA(cudaMemcpyAsync(frontier.get(), groups.get(), ng * sizeof(Key), cudaMemcpyDeviceToDevice, s));
    DevBuf<unsigned long long> counter(2, s);
    t.balance_rounds = 0;
    // membership table of the node set: built once with room to grow (balancing adds a few percent), the groups a
    // pass creates are inserted; `groups` is appended to and sorted once at the end
    KeyTable table;
    size_t table_room = ng + ng / 4 + 4096;
    if (nf > 0) table.build(groups.get(), ng, s, table_room);
    size_t groups_cap = ng;
    const size_t ng_before_balance = ng;
    while (nf > 0) {
        size_t cap = std::max<size_t>(nf * 4, 1 << 16);
        DevBuf<Key> rec_key;
        DevBuf<unsigned long long> rec_rank;
        unsigned long long produced = 0;
        for (;;) {
            rec_key.alloc(cap, s);
            rec_rank.alloc(cap, s);
            ASRB_CUDA(cudaMemsetAsync(counter.get(), 0, 2 * sizeof(unsigned long long), s));
            balance_round_kernel<<<grid_for(nf * 6, 256), 256, 0, s>>>(frontier.get(), (long long)nf, table.view(),
                                                                      rec_key.get(), rec_rank.get(), cap, counter.get());
            ASRB_CHECK_LAUNCH();
            produced = d2h_scalar(counter.get(), s);
            if (produced <= cap) break;
            cap = (size_t)produced;  // overflow: rerun the round with an exact-size buffer
        }
        ++t.balance_rounds;
        pt.lap("octree: balance emit", (long long)produced);
        if (produced == 0) break;
        // records ordered by (group, rank): stable radix sorts, rank first
        sort_pairs_u64_u64((Key*)rec_rank.get(), (unsigned long long*)rec_key.get(), (size_t)produced, s);
        sort_pairs_u64_u64(rec_key.get(), rec_rank.get(), (size_t)produced, s);
        DevBuf<Key> new_key((size_t)produced, s);
        DevBuf<unsigned long long> new_rank((size_t)produced, s);
        ASRB_CUDA(cudaMemsetAsync(counter.get(), 0, 2 * sizeof(unsigned long long), s));
        record_heads_kernel<<<grid_for(produced, 256), 256, 0, s>>>(rec_key.get(), rec_rank.get(), (long long)produced,
                                                                   new_key.get(), new_rank.get(), counter.get());
        ASRB_CHECK_LAUNCH();
        size_t nn = (size_t)d2h_scalar(counter.get(), s);
        sort_pairs_u64_u64(new_key.get(), new_rank.get(), nn, s);  // the heads were appended in arbitrary order
        const size_t cand_cap = std::max<size_t>(nn, 1);  // a candidate's child group is one of the nn new groups
        DevBuf<long long> cand(cand_cap, s);
        DevBuf<Key> cand_key(cand_cap, s);
        balance_candidates_kernel<<<grid_for(nf, 256), 256, 0, s>>>(frontier.get(), (long long)nf, table.view(),
                                                                   new_key.get(), (long long)nn, cand.get(),
                                                                   cand_key.get(), cand_cap, counter.get() + 1);
        ASRB_CHECK_LAUNCH();
        const size_t nc = (size_t)d2h_scalar(counter.get() + 1, s);
        pt.lap("octree: balance sort+cand", (long long)nc);
        if (nc > 0) {
            // sequential resolution on the host, in sweep order
            ASRB_REQUIRE(nc <= cand_cap, "octree balance: too many order-dependent keys in one pass");
            std::vector<long long> cpos_u(nc);
            std::vector<Key> ckey_u(nc);
            std::vector<Key> hk((size_t)produced);
            std::vector<unsigned long long> hr((size_t)produced);
            ASRB_CUDA(cudaMemcpyAsync(cpos_u.data(), cand.get(), nc * sizeof(long long), cudaMemcpyDeviceToHost, s));
            ASRB_CUDA(cudaMemcpyAsync(ckey_u.data(), cand_key.get(), nc * sizeof(Key), cudaMemcpyDeviceToHost, s));
            ASRB_CUDA(cudaMemcpyAsync(hk.data(), rec_key.get(), produced * sizeof(Key), cudaMemcpyDeviceToHost, s));
            ASRB_CUDA(cudaMemcpyAsync(hr.data(), rec_rank.get(), produced * sizeof(unsigned long long),
                                      cudaMemcpyDeviceToHost, s));
            ASRB_CUDA(cudaStreamSynchronize(s));
            std::vector<size_t> order(nc);
            for (size_t c = 0; c < nc; ++c) order[c] = c;
            std::sort(order.begin(), order.end(), [&](size_t x, size_t y) { return cpos_u[x] < cpos_u[y]; });
            std::vector<long long> cpos(nc);
            std::vector<Key> ckey(nc);
            for (size_t c = 0; c < nc; ++c) {
                cpos[c] = cpos_u[order[c]];
                ckey[c] = ckey_u[order[c]];
            }
            std::vector<char> invalid(nc, 0);
            auto emitter_invalid = [&](unsigned long long j) {
                auto it = std::lower_bound(cpos.begin(), cpos.end(), (long long)j);
                return it != cpos.end() && *it == (long long)j && invalid[it - cpos.begin()];
            };
            bool any_invalid = false;
            for (size_t c = 0; c < nc; ++c) {
                const Key child = ckey[c] << 3;
                size_t p = std::lower_bound(hk.begin(), hk.end(), child) - hk.begin();
                for (; p < hk.size() && hk[p] == child; ++p) {
                    const unsigned long long j = hr[p] / kRankPerKey;
                    if (j >= (unsigned long long)cpos[c]) break;  // records of a key are ordered by rank
                    if (!emitter_invalid(j)) {
                        invalid[c] = 1;
                        any_invalid = true;
                        break;
                    }
                }
            }
            if (any_invalid) {
                // withdraw the events of the skipped keys and rebuild the list of created groups
                std::vector<Key> nk;
                std::vector<unsigned long long> nr;
                for (size_t p = 0; p < hk.size(); ++p) {
                    if (emitter_invalid(hr[p] / kRankPerKey)) continue;
                    if (!nk.empty() && nk.back() == hk[p]) continue;  // (key, rank) order: the first survivor is the earliest
                    nk.push_back(hk[p]);
                    nr.push_back(hr[p]);
                }
                nn = nk.size();
                ASRB_CUDA(cudaMemcpyAsync(new_key.get(), nk.data(), nn * sizeof(Key), cudaMemcpyHostToDevice, s));
                ASRB_CUDA(cudaMemcpyAsync(new_rank.get(), nr.data(), nn * sizeof(unsigned long long),
                                          cudaMemcpyHostToDevice, s));
                ASRB_CUDA(cudaStreamSynchronize(s));
            }
        }
        pt.lap("octree: balance resolve", (long long)nn);
        if (nn == 0) break;
        // the new groups (disjoint from `groups` by construction) are appended and added to the membership table
        if (ng + nn > groups_cap) {
            const size_t new_cap = std::max(ng + nn, groups_cap + groups_cap / 8 + 4096);
            DevBuf<Key> grown(new_cap, s);
            ASRB_CUDA(cudaMemcpyAsync(grown.get(), groups.get(), ng * sizeof(Key), cudaMemcpyDeviceToDevice, s));
            groups = std::move(grown);
            groups_cap = new_cap;
        }
        ASRB_CUDA(cudaMemcpyAsync(groups.get() + ng, new_key.get(), nn * sizeof(Key), cudaMemcpyDeviceToDevice, s));
        if (ng + nn > table_room) {  // rare: the table would exceed its load factor -> rebuild with more room
            table_room = (ng + nn) + (ng + nn) / 4 + 4096;
            table.build(groups.get(), ng + nn, s, table_room);
        } else {
            table.insert(new_key.get(), nn, ng, s);
        }
        ng += nn;
        // next pass: the created groups in creation order
        sort_pairs_u64_u64((Key*)new_rank.get(), (unsigned long long*)new_key.get(), nn, s);
        frontier.alloc(nn, s);
        ASRB_CUDA(cudaMemcpyAsync(frontier.get(), new_key.get(), nn * sizeof(Key), cudaMemcpyDeviceToDevice, s));
        nf = nn;
        pt.lap("octree: balance merge", (long long)ng);
    }
    if (ng > ng_before_balance) sort_keys_u64(groups.get(), ng, s);  // ascending node order for the leaf pass

    // 4. nodes, leaf flags, sorted leaves
    t.num_groups = (int64_t)ng;
    bool group0 = false;
    if (ng) group0 = d2h_scalar(groups.get(), s) == 0;
    t.root_separate = t.any && !group0;
    t.num_nodes = (int64_t)ng * 8 + (t.root_separate ? 1 : 0);
    t.groups = std::move(groups);
    t.group_table.build(t.groups.get(), ng, s);
    t.node_leaf.alloc((size_t)t.num_nodes, s);
    t.node_rank.alloc((size_t)t.num_nodes + 1, s);
    if (t.num_nodes) {
        leaf_flag_kernel<<<grid_for(t.num_nodes, 256), 256, 0, s>>>(t.groups.get(), t.group_table.view(),
                                                                    t.root_separate, t.num_nodes, t.node_leaf.get());
        ASRB_CHECK_LAUNCH();
    }
    exclusive_sum_u8_to_i64(t.node_leaf.get(), t.node_rank.get(), (size_t)t.num_nodes, s);
    t.num_leaves = d2h_scalar(t.node_rank.get() + t.num_nodes, s);
    t.leaves.alloc((size_t)t.num_leaves, s);
    if (t.num_nodes) {
        leaf_compact_kernel<<<grid_for(t.num_nodes, 256), 256, 0, s>>>(t.groups.get(), t.root_separate, t.num_nodes,
                                                                       t.node_leaf.get(), t.node_rank.get(),
                                                                       t.leaves.get());
        ASRB_CHECK_LAUNCH();
    }
    pt.lap("octree: leaves", (long long)t.num_leaves);
}

}  // namespace asrb
