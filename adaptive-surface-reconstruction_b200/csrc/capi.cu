// extern "C" surface of libasr_b200.so — see include/asr_b200.h for the contract
// and the reference interfaces each entry point replaces.
#include "../../include/asr_b200.h"

#include "internal.h"
#include "profile.cuh"
#include "search.h"
#include "sparse_conv.h"
#include "spconv_gx.h"

namespace asrb {
std::atomic<long long> g_kernel_launches{0};

void continuous_conv(const float* filters, const float* out_pos, const float* extents, int extents_stride,
                     const float* offset, const float* inp_pos, const float* inp_feat, const float* inp_importance,
                     const int32_t* nidx, const float* nimp, const int64_t* splits, int64_t V, int S, int Cin, int Cout,
                     int normalize, const float* bias, int relu, float* out, cudaStream_t s);
void aggregation_importance(const float* compat, const float* d2, int64_t n, float* out, cudaStream_t s);
void decode_mlp(const float* shifts, const float* code, int64_t V, const float* w1, const float* b1, const float* w2,
                const float* b2, const float* w3, const float* signed_scale, float* values, float* grad,
                cudaStream_t s);
void contour_count(const float* values, const int64_t* duals, int64_t D, float thr, uint8_t* flag, int64_t* offset,
                   int64_t* num_vertices, cudaStream_t s);
void contour_fill(const float* values, const int64_t* duals, int64_t D, float thr, const float* pos,
                  const uint8_t* flag, const int64_t* offset, float* vertices, int64_t* vertex_dual, cudaStream_t s);
struct TriPlan;
TriPlan* contour_triangles_create(const float* values, const int64_t* duals, int64_t D, float thr,
                                  const int64_t* vertex_dual, int64_t M, int64_t V, int64_t* num_triangles,
                                  int64_t* num_extra, cudaStream_t s);
void contour_triangles_fill(TriPlan& P, float* vertices, int32_t* triangles, cudaStream_t s);
void contour_triangles_destroy(TriPlan* P);
void mesh_components(const int32_t* triangles, int64_t T, int64_t V, int64_t* label, int64_t* size, cudaStream_t s);
size_t packed_weights_floats(int K, int N);
void pack_weights(const float* W, int K, int N, float* out, cudaStream_t s);
void dense_gemm_tf32x3(const float* A, int64_t M, int K, int lda, const float* Wp, int N, const float* bias, int relu,
                       float* D, int ldd, cudaStream_t s);
void shard_positions(const Key* keys, int64_t V, unsigned long long* pos, cudaStream_t s);
void shard_owner(const Key* keys, int64_t V, const unsigned long long* thr, int nthr, uint8_t* owner, cudaStream_t s);
void shard_need_mask(const int64_t* splits, const int32_t* idx, int64_t V_out, const uint8_t* owner_out,
                     const uint8_t* owner_in, int me, uint32_t* mask, cudaStream_t s);
void shard_push(void* const* peer_base, int world, int me, int64_t offset, int64_t pitch, const int* seg_off, int nseg,
                int seg_len, const int32_t* rows, int64_t nrows, const uint32_t* mask, cudaStream_t s);
void profile_set(bool on);
void profile_reset();
int profile_count();
bool profile_get(int i, std::string& name, double& ms, long long& launches, double& flops);
void invert_neighbors_list(int64_t num_points, const int32_t* idx, const int64_t* splits, int64_t Q, int64_t E,
                           const void* attrs, int attr_bytes, int32_t* out_idx, int64_t* out_splits, void* out_attrs,
                           cudaStream_t s);
}  // namespace asrb

using namespace asrb;

struct asr_octree {
    Octree t;
};
struct asr_search {
    Search s;
};
struct asr_conv_plan {
    ConvPlan p;
};
struct asr_gx_plan {
    gx::Plan p;
};

namespace {
thread_local std::string g_error;

template <class F>
int guarded(F&& f) {
    try {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
            throw Error(kCudaError, "asr_b200: no CUDA device available (there is no CPU fallback)");
        f();
        return kOk;
    } catch (const Error& e) {
        g_error = e.what();
        return e.code;
    } catch (const std::exception& e) {
        g_error = e.what();
        return kRuntimeError;
    }
}
inline cudaStream_t S(void* stream) { return (cudaStream_t)stream; }

template <class T>
void copy_out(T* dst, const DevBuf<T>& src, size_t n, cudaStream_t s) {
    if (dst && n) ASRB_CUDA(cudaMemcpyAsync(dst, src.get(), n * sizeof(T), cudaMemcpyDeviceToDevice, s));
}

__global__ void iota_i64_kernel(int64_t* out, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}
}  // namespace

extern "C" {

int asr_version(void) { return ASR_B200_VERSION; }
const char* asr_last_error(void) { return g_error.c_str(); }
int64_t asr_kernel_launches(void) { return (int64_t)g_kernel_launches.load(); }

int asr_set_option(const char* name, int value) {
    return guarded([&] {
        ASRB_REQUIRE(name != nullptr, "option name is null");
        if (std::string(name) == "conv_row_block_shift") sparse_conv_row_block_shift(value);
        else if (std::string(name) == "tc_ntile") sparse_conv_tc_ntile(value);
        else if (std::string(name) == "gx_acc_groups") gx::set_acc_groups(value);
        else if (std::string(name) == "gx_tma_gather") gx::set_tma_gather(value);
        else if (std::string(name) == "gx_l1_gather") gx::set_l1_gather(value);
        else if (std::string(name) == "gx_max_stages") gx::set_max_stages(value);
        else if (std::string(name) == "gx_ablate") gx::set_ablate(value);
        else if (std::string(name) == "gx_single_tmem") gx::set_single_tmem(value);
        else if (std::string(name) == "gx_one_team") gx::set_one_team(value);
        else if (std::string(name) == "gx_trace") gx::set_trace(value);
        else if (std::string(name) == "tc_stages") sparse_conv_tc_tune(value, 0);
        else if (std::string(name) == "tc_row_groups") sparse_conv_tc_tune(0, value);
        else throw Error(kInvalidArgument, std::string("unknown option: ") + name);
    });
}

int asr_pool_stats(int64_t* reserved_bytes, int64_t* used_bytes, int64_t* release_threshold, int64_t* used_high_bytes) {
    return guarded([&] {
        int dev = 0;
        ASRB_CUDA(cudaGetDevice(&dev));
        cudaMemPool_t pool;
        ASRB_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
        uint64_t v = 0;
        if (reserved_bytes) {
            ASRB_CUDA(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &v));
            *reserved_bytes = (int64_t)v;
        }
        if (used_bytes) {
            ASRB_CUDA(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &v));
            *used_bytes = (int64_t)v;
        }
        if (used_high_bytes) {
            ASRB_CUDA(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemHigh, &v));
            *used_high_bytes = (int64_t)v;
        }
        if (release_threshold) {
            ASRB_CUDA(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &v));
            *release_threshold = v > (uint64_t)INT64_MAX ? INT64_MAX : (int64_t)v;
        }
    });
}

void asr_profile_enable(int on) { profile_set(on != 0); }
void asr_profile_reset(void) { profile_reset(); }
int asr_profile_count(void) { return profile_count(); }
int asr_profile_get(int i, char* name, int name_cap, double* total_ms, int64_t* launches, double* flops) {
    std::string n;
    double ms = 0, fl = 0;
    long long l = 0;
    if (!profile_get(i, n, ms, l, fl)) return 1;
    if (name && name_cap > 0) {
        snprintf(name, (size_t)name_cap, "%s", n.c_str());
    }
    if (total_ms) *total_ms = ms;
    if (launches) *launches = l;
    if (flops) *flops = fl;
    return 0;
}

int asr_octree_create(const float* d_points, const float* d_radii, int64_t num_points, const float h_bb_min[3],
                      const float h_bb_max[3], float radius_scale, int grow_steps, int max_depth, void* stream,
                      asr_octree** out) {
    return guarded([&] {
        ASRB_REQUIRE(out != nullptr, "out must not be null");
        ASRB_REQUIRE(num_points >= 0, "num_points must be >= 0");
        ASRB_REQUIRE(grow_steps == 0, "grow_steps must be 0 (the only value the reference pipeline uses)");
        ASRB_REQUIRE(num_points == 0 || (d_points && d_radii), "points/radii must not be null");
        auto h = std::make_unique<asr_octree>();
        h->t.frame = make_frame(h_bb_min, h_bb_max);
        cudaGetDevice(&h->t.device);
        octree_build(h->t, d_points, d_radii, num_points, radius_scale, max_depth, S(stream));
        *out = h.release();
    });
}
void asr_octree_destroy(asr_octree* tree) { delete tree; }
int64_t asr_octree_num_leaves(const asr_octree* tree) { return tree ? tree->t.num_leaves : 0; }
int64_t asr_octree_num_nodes(const asr_octree* tree) { return tree ? tree->t.num_nodes : 0; }
int asr_octree_balance_rounds(const asr_octree* tree) { return tree ? tree->t.balance_rounds : 0; }
int asr_octree_get_leaves(const asr_octree* tree, uint64_t* d_out, void* stream) {
    return guarded([&] {
        ASRB_REQUIRE(tree, "tree is null");
        copy_out((Key*)d_out, tree->t.leaves, (size_t)tree->t.num_leaves, S(stream));
    });
}
int asr_octree_get_frame(const asr_octree* tree, float* vs, float* ivs, int32_t* off) {
    return guarded([&] {
        ASRB_REQUIRE(tree, "tree is null");
        for (int l = 0; l <= kMaxLevel; ++l) {
            if (vs) vs[l] = tree->t.frame.vs[l];
            if (ivs) ivs[l] = tree->t.frame.ivs[l];
        }
        if (off)
            for (int a = 0; a < 3; ++a) off[a] = tree->t.frame.off[a];
    });
}

int asr_grids_build(asr_octree* tree, int num_levels, int voxel_info_all_levels, void* stream) {
    return guarded([&] {
        ASRB_REQUIRE(tree, "tree is null");
        grids_build(tree->t, num_levels, voxel_info_all_levels != 0, S(stream));
    });
}
int asr_grids_level_size(const asr_octree* tree, int level, int64_t* num_voxels, int64_t* num_neighbors) {
    return guarded([&] {
        ASRB_REQUIRE(tree, "tree is null");
        ASRB_REQUIRE(level >= 0 && level < (int)tree->t.grids.size(), "level out of range (call asr_grids_build first)");
        if (num_voxels) *num_voxels = tree->t.grids[level]->V;
        if (num_neighbors) *num_neighbors = tree->t.grids[level]->E;
    });
}
int asr_grids_get(const asr_octree* tree, int level, uint64_t* keys, float* centers, float* sizes, int32_t* nidx,
                  uint8_t* nslot, int64_t* nsplits, int32_t* uidx, uint8_t* uslot, int64_t* usplits, void* stream) {
    return guarded([&] {
        ASRB_REQUIRE(tree, "tree is null");
        ASRB_REQUIRE(level >= 0 && level < (int)tree->t.grids.size(), "level out of range (call asr_grids_build first)");
        const GridLevel& g = *tree->t.grids[level];
        cudaStream_t s = S(stream);
        const size_t V = (size_t)g.V, E = (size_t)g.E;
        copy_out((Key*)keys, g.keys, V, s);
        if (centers || sizes) ASRB_REQUIRE(g.centers.size() == 3 * V, "voxel info was not built for this level");
        copy_out(centers, g.centers, 3 * V, s);
        copy_out(sizes, g.sizes, V, s);
        copy_out(nidx, g.nidx, E, s);
        copy_out(nslot, g.nslot, E, s);
        copy_out(nsplits, g.nsplits, V + 1, s);
        if (uidx || uslot || usplits) ASRB_REQUIRE(g.has_up, "the last level has no up table");
        copy_out(uidx, g.uidx, V, s);
        copy_out(uslot, g.uslot, V, s);
        if (usplits) {
            iota_i64_kernel<<<grid_for(V + 1, 256), 256, 0, s>>>(usplits, (long long)V + 1);
            ASRB_CHECK_LAUNCH();
        }
    });
}

int asr_duals_count(asr_octree* tree, int64_t* num_duals, void* stream) {
    return guarded([&] {
        ASRB_REQUIRE(tree && num_duals, "null argument");
        duals_count(tree->t, S(stream));
        *num_duals = tree->t.num_duals;
    });
}
int asr_duals_begin(asr_octree* tree, void* stream) {
    return guarded([&] {
        ASRB_REQUIRE(tree, "tree is null");
        duals_begin(tree->t, S(stream));
    });
}
int asr_duals_fill(asr_octree* tree, int64_t* d_out, void* stream) {
    return guarded([&] {
        ASRB_REQUIRE(tree, "tree is null");
        duals_fill(tree->t, d_out, S(stream));
    });
}
int asr_duals_check(asr_octree* tree) {
    return guarded([&] {
        ASRB_REQUIRE(tree, "tree is null");
        duals_check(tree->t);
    });
}

int asr_radius_search_create(const float* d_points, int64_t num_points, const float* d_queries, const float* d_radii,
                             int64_t num_queries, const float* h_frame, void* stream, asr_search** out,
                             int64_t* num_pairs) {
    return guarded([&] {
        ASRB_REQUIRE(out && num_pairs, "null argument");
        ASRB_REQUIRE(num_points >= 0 && num_queries >= 0, "negative size");
        ASRB_REQUIRE(num_points < (int64_t(1) << 31), "too many points");
        auto h = std::make_unique<asr_search>();
        ASRB_REQUIRE(!h_frame || h_frame[3] > 0.f, "frame cell size must be positive");
        search_prepare(h->s, d_points, num_points, d_queries, d_radii, num_queries, h_frame, S(stream));
        *num_pairs = h->s.num_pairs;
        *out = h.release();
    });
}
int asr_radius_search_fill(asr_search* search, int32_t* idx, float* dist, int64_t* splits, void* stream) {
    return guarded([&] {
        ASRB_REQUIRE(search && splits, "null argument");
        search_fill(search->s, idx, dist, splits, S(stream));
    });
}
void asr_radius_search_destroy(asr_search* search) { delete search; }

int asr_kdtree_create(const float* d_points, int64_t num_points, void* stream, asr_search** out) {
    return guarded([&] {
        ASRB_REQUIRE(out, "null argument");
        ASRB_REQUIRE(num_points >= 0 && num_points < (int64_t(1) << 31), "bad point count");
        auto h = std::make_unique<asr_search>();
        knn_build(h->s, d_points, num_points, S(stream));
        *out = h.release();
    });
}
int asr_kdtree_k_radius(asr_search* tree, int k, float* d_out, void* stream) {
    return guarded([&] {
        ASRB_REQUIRE(tree, "null handle");
        knn_radius(tree->s, k, d_out, S(stream));
    });
}
int asr_kdtree_inlier(asr_search* tree, const float* d_radii, float radius_fraction, int k, int outlier_threshold,
                      uint8_t* d_out, void* stream) {
    return guarded([&] {
        ASRB_REQUIRE(tree, "null handle");
        knn_inlier(tree->s, d_radii, radius_fraction, k, outlier_threshold, d_out, S(stream));
    });
}
int asr_radius_neighbor_counts(const float* d_points, int64_t num_points, const float* d_radii, int32_t* d_out,
                               void* stream) {
    return guarded([&] { radius_neighbor_counts(d_points, num_points, d_radii, d_out, S(stream)); });
}

int asr_scale_compatibility(const float* sizes, const float* radii, const int32_t* idx, const int64_t* splits,
                            int64_t num_queries, float* out, void* stream) {
    return guarded([&] { scale_compat(sizes, radii, idx, splits, num_queries, out, S(stream)); });
}
int asr_aggregation_importance(const float* compat, const float* dist, int64_t n, float* out, void* stream) {
    return guarded([&] { aggregation_importance(compat, dist, n, out, S(stream)); });
}

int asr_continuous_conv(const float* filters, const float* out_pos, const float* extents, int extents_stride,
                        const float* offset, const float* inp_pos, const float* inp_feat, const float* inp_importance,
                        const int32_t* nidx, const float* nimp, const int64_t* splits, int64_t num_out, int kernel_size,
                        int in_channels, int out_channels, int normalize, const float* bias, int relu, float* out,
                        void* stream) {
    return guarded([&] {
        ASRB_REQUIRE(kernel_size >= 1 && in_channels >= 1 && out_channels >= 1, "bad filter shape");
        ASRB_REQUIRE(extents_stride == 0 || extents_stride == 1, "extents_stride must be 0 or 1");
        continuous_conv(filters, out_pos, extents, extents_stride, offset, inp_pos, inp_feat, inp_importance, nidx, nimp,
                        splits, num_out, kernel_size, in_channels, out_channels, normalize, bias, relu, out, S(stream));
    });
}

int asr_conv_plan_create(const int32_t* idx, const uint8_t* slot, const int64_t* splits, int64_t num_out,
                         int64_t num_entries, int kernel_size, void* stream, asr_conv_plan** out) {
    return guarded([&] {
        ASRB_REQUIRE(out, "null argument");
        auto h = std::make_unique<asr_conv_plan>();
        conv_plan_build(h->p, idx, slot, splits, num_out, num_entries, kernel_size, S(stream));
        *out = h.release();
    });
}
void asr_conv_plan_destroy(asr_conv_plan* plan) { delete plan; }

int64_t asr_packed_conv_filters_size(int kernel_size, int in_channels, int out_channels) {
    return (int64_t)packed_conv_filters_floats(kernel_size, in_channels, out_channels);
}
int asr_pack_conv_filters(const float* d_filters, int kernel_size, int in_channels, int out_channels, float* d_packed,
                          void* stream) {
    return guarded([&] {
        ASRB_REQUIRE(kernel_size >= 1 && in_channels >= 1 && out_channels >= 1 && out_channels <= 256,
                     "pack_conv_filters: bad shape (out_channels must be <= 256)");
        pack_conv_filters(d_filters, kernel_size, in_channels, out_channels, d_packed, S(stream));
    });
}

int asr_sparse_conv(const asr_conv_plan* plan, const float* filters, const float* packed_filters, const float* x,
                    int in_channels, int out_channels,
                    const float* inp_importance, const float* neighbors_importance, int importance_col, int normalize,
                    int normalize_col, const float* normalizer, const int64_t* splits, const float* bias, int relu,
                    float* out, void* stream) {
    return guarded([&] {
        ASRB_REQUIRE(plan, "plan is null");
        ASRB_REQUIRE(!normalize || normalizer || splits, "normalize needs a normalizer or the row splits");
        ASRB_REQUIRE(filters || packed_filters, "filters is null");
        sparse_conv_forward(plan->p, x, filters, packed_filters, in_channels, out_channels, inp_importance,
                            neighbors_importance,
                            importance_col, normalize, normalize_col, normalizer, splits, bias, relu, out, S(stream));
    });
}
int asr_gx_plan_begin(const int32_t* idx, const uint8_t* slot, const int64_t* splits, int64_t num_out, int64_t num_in,
                      int64_t num_entries, int kernel_size, int mode, const int32_t* row_map, void* stream,
                      asr_gx_plan** out) {
    return guarded([&] {
        ASRB_REQUIRE(out, "null argument");
        ASRB_REQUIRE(num_out >= 0 && num_in >= 0 && num_entries >= 0, "negative size");
        auto h = std::make_unique<asr_gx_plan>();
        gx::plan_begin(h->p, idx, slot, splits, num_out, num_in, num_entries, kernel_size, mode, row_map, S(stream));
        *out = h.release();
    });
}
int asr_gx_plan_finish(asr_gx_plan* plan, void* stream, int64_t* num_rare) {
    return guarded([&] {
        ASRB_REQUIRE(plan, "plan is null");
        gx::plan_finish(plan->p, S(stream));
        if (num_rare) *num_rare = plan->p.mode == gx::kModeStationary ? plan->p.R : 0;
    });
}
void asr_gx_plan_destroy(asr_gx_plan* plan) { delete plan; }
int64_t asr_gx_packed_filters_bytes(int kernel_size, int in_channels, int ncols) {
    return (int64_t)gx::packed_filter_bytes(kernel_size, in_channels, ncols);
}
int asr_gx_pack_filters(const float* d_filters, int kernel_size, int in_channels, int out_channels, int col0, int ncols,
                        int scale_exp, void* d_packed, void* stream) {
    return guarded([&] {
        ASRB_REQUIRE(d_filters && d_packed && kernel_size >= 1, "gx pack: null argument");
        gx::pack_filters(d_filters, kernel_size, in_channels, out_channels, col0, ncols, scale_exp, d_packed, S(stream));
    });
}
static gx::H2View h2view(const void* p, int64_t num_rows, int C, int pitch, int hi, int lo) {
    gx::H2View v;
    v.p = (__half*)p;
    v.rows = num_rows + 1;
    v.C = C;
    v.pitch = pitch;
    v.hi = hi;
    v.lo = lo;
    return v;
}
int asr_gx_from_f32(const float* d_x, int64_t num_rows, int channels, int ldx, const float* d_row_scale,
                    const int32_t* d_rows, int64_t out_rows, void* d_out, int out_pitch, int out_hi, int out_lo,
                    void* stream) {
    return guarded([&] {
        gx::from_f32(d_x, num_rows, channels, ldx, d_row_scale, d_rows,
                     h2view(d_out, d_rows ? out_rows : num_rows, channels, out_pitch, out_hi, out_lo), S(stream));
    });
}
int asr_gx_to_f32(const void* d_x, int64_t num_rows, int channels, int pitch, int hi, int lo, float* d_out, int ldo,
                  void* stream) {
    return guarded([&] { gx::to_f32(h2view(d_x, num_rows, channels, pitch, hi, lo), num_rows, d_out, ldo, S(stream)); });
}
int asr_gx_scale_rows(const void* d_x, int64_t num_rows, int channels, int pitch, int hi, int lo,
                      const float* d_row_scale, void* d_out, int out_pitch, int out_hi, int out_lo, void* stream) {
    return guarded([&] {
        ASRB_REQUIRE(d_row_scale, "gx scale_rows: row_scale is null");
        gx::scale_rows(h2view(d_x, num_rows, channels, pitch, hi, lo), num_rows, d_row_scale,
                       h2view(d_out, num_rows, channels, out_pitch, out_hi, out_lo), S(stream));
    });
}
int asr_gx_conv(const asr_gx_plan* plan, const void* d_x, int in_channels, int x_pitch, int x_hi, int x_lo,
                const void* d_packed, int ncols, int scale_exp, const float* d_bias, int relu, const float* d_norm,
                const float* d_imp, const void* d_res, int res_pitch, int res_hi, int res_lo, void* d_out_h2, int out_pitch, int out_hi,
                int out_lo, float* d_out_f32, int out_f32_pitch, float* d_pairbuf, void* stream) {
    return guarded([&] {
        ASRB_REQUIRE(plan && d_x && d_packed, "gx conv: null argument");
        gx::ConvArgs a;
        a.x = h2view(d_x, plan->p.V_in, in_channels, x_pitch, x_hi, x_lo);
        a.wp = d_packed;
        a.K = plan->p.K;
        a.ncols = ncols;
        a.scale_exp = scale_exp;
        a.bias = d_bias;
        a.relu = relu;
        a.norm = d_norm;
        a.imp = d_imp;
        if (d_res) a.res = h2view(d_res, plan->p.V, ncols, res_pitch, res_hi, res_lo);
        if (d_out_h2) a.out = h2view(d_out_h2, plan->p.V, ncols, out_pitch, out_hi, out_lo);
        a.out_f32 = d_out_f32;
        a.out_f32_pitch = out_f32_pitch;
        a.pairbuf = d_pairbuf;
        gx::conv(plan->p, a, S(stream));
    });
}
int asr_shard_positions(const uint64_t* d_keys, int64_t num_voxels, uint64_t* d_pos, void* stream) {
    return guarded([&] { shard_positions((const Key*)d_keys, num_voxels, (unsigned long long*)d_pos, S(stream)); });
}
int asr_shard_owner(const uint64_t* d_keys, int64_t num_voxels, const uint64_t* d_thresholds, int num_thresholds,
                    uint8_t* d_owner, void* stream) {
    return guarded([&] {
        shard_owner((const Key*)d_keys, num_voxels, (const unsigned long long*)d_thresholds, num_thresholds, d_owner,
                    S(stream));
    });
}
int asr_shard_need_mask(const int64_t* d_row_splits, const int32_t* d_index, int64_t num_out, const uint8_t* d_owner_out,
                        const uint8_t* d_owner_in, int rank, uint32_t* d_mask, void* stream) {
    return guarded([&] { shard_need_mask(d_row_splits, d_index, num_out, d_owner_out, d_owner_in, rank, d_mask, S(stream)); });
}
int asr_shard_push(void* const* peer_base, int world, int rank, int64_t offset, int64_t pitch, const int* seg_off, int nseg,
                   int seg_len, const int32_t* d_rows, int64_t num_rows, const uint32_t* d_mask, void* stream) {
    return guarded([&] {
        ASRB_REQUIRE(peer_base && seg_off, "shard_push: null argument");
        shard_push(peer_base, world, rank, offset, pitch, seg_off, nseg, seg_len, d_rows, num_rows, d_mask, S(stream));
    });
}

int asr_gx_trace(void* stream, int ctas, unsigned* counters) {
    return guarded([&] {
        ASRB_REQUIRE(counters != nullptr && ctas > 0, "asr_gx_trace: null output");
        gx::trace_read(counters, ctas, S(stream));
    });
}

int asr_gx_overflow(void* stream, int* flag) {
    return guarded([&] {
        ASRB_REQUIRE(flag, "null argument");
        *flag = gx::overflow_flag_read_and_clear(S(stream));
    });
}

int asr_reduce_subarrays_sum(const float* values, const int32_t* index, const int64_t* splits, int64_t num_rows,
                             float* out, void* stream) {
    return guarded([&] { row_importance(values, index, splits, num_rows, out, S(stream)); });
}
int asr_invert_neighbors_list(int64_t num_points, const int32_t* idx, const int64_t* splits, int64_t num_queries,
                              int64_t num_entries, const void* attrs, int attr_bytes, int32_t* out_idx,
                              int64_t* out_splits, void* out_attrs, void* stream) {
    return guarded([&] {
        ASRB_REQUIRE(num_points >= 0 && num_entries >= 0, "negative size");
        invert_neighbors_list(num_points, idx, splits, num_queries, num_entries, attrs, attr_bytes, out_idx, out_splits,
                              out_attrs, S(stream));
    });
}

int asr_decode(const float* shifts, const float* code, int64_t V, const float* w1, const float* b1, const float* w2,
               const float* b2, const float* w3, const float* signed_scale, float* values, float* grad, void* stream) {
    return guarded([&] { decode_mlp(shifts, code, V, w1, b1, w2, b2, w3, signed_scale, values, grad, S(stream)); });
}

int64_t asr_packed_weights_size(int in_features, int out_features) {
    return (int64_t)packed_weights_floats(in_features, out_features);
}
int asr_pack_weights(const float* d_w, int in_features, int out_features, float* d_packed, void* stream) {
    return guarded([&] {
        ASRB_REQUIRE(in_features >= 1 && out_features >= 1 && out_features <= 256, "bad weight shape");
        pack_weights(d_w, in_features, out_features, d_packed, S(stream));
    });
}
int asr_dense_tf32x3(const float* d_a, int64_t rows, int in_features, int lda, const float* d_packed_w, int out_features,
                     const float* d_bias, int relu, float* d_out, int ldd, void* stream) {
    return guarded([&] {
        ASRB_REQUIRE(rows >= 0 && in_features >= 1, "bad shape");
        ASRB_REQUIRE(lda >= in_features && ldd >= out_features, "leading dimensions too small");
        dense_gemm_tf32x3(d_a, rows, in_features, lda, d_packed_w, out_features, d_bias, relu, d_out, ldd, S(stream));
    });
}

int asr_contour_count(const float* values, const int64_t* duals, int64_t D, float thr, uint8_t* flag, int64_t* offset,
                      int64_t* num_vertices, void* stream) {
    return guarded([&] {
        ASRB_REQUIRE(num_vertices, "null argument");
        contour_count(values, duals, D, thr, flag, offset, num_vertices, S(stream));
    });
}
int asr_contour_fill(const float* values, const int64_t* duals, int64_t D, float thr, const float* pos,
                     const uint8_t* flag, const int64_t* offset, float* vertices, int64_t* vertex_dual, void* stream) {
    return guarded([&] { contour_fill(values, duals, D, thr, pos, flag, offset, vertices, vertex_dual, S(stream)); });
}

int asr_contour_triangles_create(const float* values, const int64_t* duals, int64_t D, float thr,
                                 const int64_t* vertex_dual, int64_t M, int64_t num_nodes, void* stream, void** handle,
                                 int64_t* num_triangles, int64_t* num_extra_vertices) {
    return guarded([&] {
        ASRB_REQUIRE(handle && num_triangles && num_extra_vertices, "null argument");
        ASRB_REQUIRE(D >= 0 && M >= 0 && num_nodes >= 0, "negative size");
        *handle = contour_triangles_create(values, duals, D, thr, vertex_dual, M, num_nodes, num_triangles,
                                           num_extra_vertices, S(stream));
    });
}
int asr_contour_triangles_fill(void* handle, float* vertices, int32_t* triangles, void* stream) {
    return guarded([&] {
        ASRB_REQUIRE(handle, "null handle");
        contour_triangles_fill(*static_cast<TriPlan*>(handle), vertices, triangles, S(stream));
    });
}
void asr_contour_triangles_destroy(void* handle) { contour_triangles_destroy(static_cast<TriPlan*>(handle)); }

int asr_mesh_components(const int32_t* triangles, int64_t num_triangles, int64_t num_vertices, int64_t* label,
                        int64_t* size, void* stream) {
    return guarded([&] {
        ASRB_REQUIRE(num_triangles >= 0 && num_vertices >= 0, "negative size");
        ASRB_REQUIRE(label || num_vertices == 0, "null argument");
        mesh_components(triangles, num_triangles, num_vertices, label, size, S(stream));
    });
}

}  // extern "C"
