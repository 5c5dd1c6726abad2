/*
 * asr_b200 — C ABI of the B200-native octree-conv -> SDF hot path.
 *
 * Drop-in boundary for the hot path of isl-org/adaptive-surface-reconstruction
 * (SURVEY.md §8b).  The reference has no FFI for this path (its only C-linkage
 * hook is the unused ASR_API_EXTERN_C macro, cpp/lib/asr_config.h:18-28); the
 * entry points below are what a binding for the reference's own interfaces
 * would call.  Each one cites the reference interface it replaces.
 *
 * Conventions
 *   - every pointer named d_* is a DEVICE pointer on the current CUDA device,
 *     h_* is a HOST pointer; buffers are owned by the caller (PyTorch);
 *   - variable-length results are two-phase: a *_count / *_create call returns the
 *     sizes, the caller allocates, a *_fill / *_get call writes;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it.  A
 *     call that returns a size synchronises that stream once;
 *   - return value: 0 ok, 1 invalid argument (reference: std::invalid_argument ->
 *     ValueError), 2 runtime error (std::runtime_error -> RuntimeError), 3 CUDA
 *     error; asr_last_error() returns the thread-local message;
 *   - there is NO CPU fallback: without a CUDA device every call fails with 3.
 */
#ifndef ASR_B200_H
#define ASR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ASR_B200_VERSION 100

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

int asr_version(void);
const char* asr_last_error(void);
/* number of kernels this library has launched since load (bench.py gpu_launches) */
int64_t asr_kernel_launches(void);

/* Development options (defaults are what the benchmarks and tests run; unknown names -> status 1).  No reference
 * counterpart.  gx convolution kernel (csrc/spconv_gx.cu): "gx_acc_groups" (accumulator groups per TMEM buffer, 0 = as
 * many as fit), "gx_single_tmem" (one TMEM buffer with two groups for 128-column shapes: half the rounding error, 2-10 %
 * slower), "gx_one_team", "gx_tma_gather" (TMA tile::gather4 instead of cp.async row gathers), "gx_l1_gather",
 * "gx_max_stages", "gx_ablate" (timing experiments only: results become garbage), "gx_trace" (instrumented build only);
 * round-1 kernels: "conv_row_block_shift", "tc_ntile", "tc_stages", "tc_row_groups". */
int asr_set_option(const char* name, int value);

/* bytes reserved / in use / release threshold / high-water mark in use of the stream-ordered memory pool the
 * library allocates from (any pointer may be NULL) */
int asr_pool_stats(int64_t* reserved_bytes, int64_t* used_bytes, int64_t* release_threshold, int64_t* used_high_bytes);
/* Per-kernel device timing (CUDA events on the launching stream) for bench.py's
 * roofline figures: enable, run, then read (name, total ms, launches, algorithmic
 * flops) per instrumented kernel.  Reading synchronises the recorded events. */
void asr_profile_enable(int on);
void asr_profile_reset(void);
int asr_profile_count(void);
int asr_profile_get(int i, char* name, int name_cap, double* total_ms, int64_t* launches, double* flops);

/* ---------------------------------------------------------------- octree
 * replaces asr::CreateOctreeFromPoints, cpp/lib/octree.h:145 / octree.cpp:230-280
 * (python: create_octree, cpp/pybind/module.cpp:372-374).  The handle mirrors the
 * opaque `Octree` object of the python module (module.cpp:282). */
typedef struct asr_octree asr_octree;

int asr_octree_create(const float* d_points, const float* d_radii, int64_t num_points,
                      const float h_bb_min[3], const float h_bb_max[3], float radius_scale,
                      int grow_steps, int max_depth, void* stream, asr_octree** out);
void asr_octree_destroy(asr_octree* tree);
int64_t asr_octree_num_leaves(const asr_octree* tree);
int64_t asr_octree_num_nodes(const asr_octree* tree);
int asr_octree_balance_rounds(const asr_octree* tree);
/* Octree::leaves (octree.h:127): sorted location codes, u64[num_leaves] */
int asr_octree_get_leaves(const asr_octree* tree, uint64_t* d_out, void* stream);
/* Octree::voxel_size / inv_voxel_size / offset (octree.h:120,131-132): host arrays [22],[22],[3] */
int asr_octree_get_frame(const asr_octree* tree, float* h_voxel_size, float* h_inv_voxel_size,
                         int32_t* h_offset);

/* ---------------------------------------------------------------- grid hierarchy
 * replaces asr::CreateGridsFromOctree, cpp/lib/grid.h:23 / grid.cpp:245-314
 * (python: create_grids_from_octree, module.cpp:402-403).  Builds `num_levels`
 * grids inside the handle; sizes per level are then queried and the arrays of
 * ASRGrid (cpp/lib/asr_types.h:30-66) copied out.  Any output pointer may be NULL.
 * voxel_* arrays exist on level 0, and on all levels iff voxel_info_all_levels;
 * up_* arrays exist on every level but the last (up row splits are arange). */
int asr_grids_build(asr_octree* tree, int num_levels, int voxel_info_all_levels, void* stream);
int asr_grids_level_size(const asr_octree* tree, int level, int64_t* num_voxels, int64_t* num_neighbors);
int asr_grids_get(const asr_octree* tree, int level, uint64_t* d_voxel_keys, float* d_voxel_centers,
                  float* d_voxel_sizes, int32_t* d_neighbors_index, uint8_t* d_neighbors_kernel_index,
                  int64_t* d_neighbors_row_splits, int32_t* d_up_neighbors_index,
                  uint8_t* d_up_neighbors_kernel_index, int64_t* d_up_neighbors_row_splits, void* stream);

/* ---------------------------------------------------------------- dual cells
 * replaces asr::CreateDualVertexIndices, cpp/lib/grid.h:28 / grid.cpp:450-459
 * (python: create_dual_vertex_indices, module.cpp:443): [num_duals, 8] leaf indices. */
int asr_duals_count(asr_octree* tree, int64_t* num_duals, void* stream);
int asr_duals_fill(asr_octree* tree, int64_t* d_dual_vertex_indices, void* stream);
/* asynchronous form: _begin queues the counting pass on `stream` without a host synchronisation (a later
 * asr_duals_count only waits for it); asr_duals_fill queues the fill; _check waits for the fill's error flag and
 * returns status 2 ("found node is not a leaf", grid.cpp:436-440) — call it before trusting the result */
int asr_duals_begin(asr_octree* tree, void* stream);
int asr_duals_check(asr_octree* tree);

/* ---------------------------------------------------------------- aggregation neighbours
 * replaces the Open3D MultiRadiusIndex/MultiRadiusSearch calls of
 * asr::ComputeAggregationNeighborsAndScaleCompatibility, cpp/lib/nsearch.h:72-80 /
 * nsearch.cpp:107-162 (python twin: o3d.core.nns multi_radius_search,
 * models/v0/datareader.py:776-785): per query all points with d^2 < r^2, ascending
 * by d^2 (ties by index); distances are SQUARED.
 * h_frame (may be NULL): {origin x, y, z, finest cell size} of the 2^21-per-axis binning grid;
 * passing the octree's frame (asr_octree_get_frame: -offset*voxel_size[21], voxel_size[21]) makes
 * the bins coincide with the voxels being queried (fewer candidates); results do not depend on it. */
typedef struct asr_search asr_search;
int asr_radius_search_create(const float* d_points, int64_t num_points, const float* d_queries,
                             const float* d_radii, int64_t num_queries, const float* h_frame, void* stream,
                             asr_search** out, int64_t* num_pairs);
int asr_radius_search_fill(asr_search* search, int32_t* d_neighbors_index, float* d_neighbors_dist,
                           int64_t* d_neighbors_row_splits, void* stream);
void asr_radius_search_destroy(asr_search* search);
/* (min(s_v, 2 r_p) / max(s_v, 2 r_p))^2 per pair, nsearch.cpp:149-161 / models/common.py:19-44 */
int asr_scale_compatibility(const float* d_voxel_sizes, const float* d_point_radii, const int32_t* d_neighbors_index,
                            const int64_t* d_neighbors_row_splits, int64_t num_queries, float* d_out, void* stream);
/* scale_compat * clamp((1 - d2)^3, 0, 1), net_definitions_torch.py:107 + common_torch.py:21-22 */
int asr_aggregation_importance(const float* d_scale_compat, const float* d_dist, int64_t num_pairs, float* d_out,
                               void* stream);

/* ---------------------------------------------------------------- continuous conv
 * replaces open3d.ml.torch.ops.continuous_conv as configured by the reference
 * (net_definitions_torch.py:53-70,108-116): ball_to_cube_radial, align_corners,
 * linear interpolation.  filters [S,S,S,Cin,Cout]; extents [V] (stride 1) or [1]
 * (stride 0); offset [3] or NULL; importance pointers may be NULL; bias may be
 * NULL; relu != 0 fuses the layer activation. */
int asr_continuous_conv(const float* d_filters, const float* d_out_positions, const float* d_extents,
                        int extents_stride, const float* d_offset, const float* d_inp_positions,
                        const float* d_inp_features, const float* d_inp_importance, const int32_t* d_neighbors_index,
                        const float* d_neighbors_importance, const int64_t* d_neighbors_row_splits,
                        int64_t num_out, int kernel_size, int in_channels, int out_channels, int normalize,
                        const float* d_bias, int relu, float* d_out_features, void* stream);

/* ---------------------------------------------------------------- generalized sparse conv
 * replaces open3d.ml.torch.ops.sparse_conv + ops.reduce_subarrays_sum as used by
 * SpecialSparseConv.forward, models/common_torch.py:95-148.  A plan is the
 * slot-sorted form of one neighbour table (neighbors_index, neighbors_kernel_index,
 * neighbors_row_splits) and is reused by every convolution on that table. */
typedef struct asr_conv_plan asr_conv_plan;
int asr_conv_plan_create(const int32_t* d_neighbors_index, const uint8_t* d_neighbors_kernel_index,
                         const int64_t* d_neighbors_row_splits, int64_t num_out, int64_t num_entries,
                         int kernel_size, void* stream, asr_conv_plan** out);
void asr_conv_plan_destroy(asr_conv_plan* plan);
/* out[o] = sum_n imp_n * x[idx_n] @ filters[slot_n]; channels >= importance_col are
 * weighted by imp_n = d_inp_importance[idx_n] (and/or d_neighbors_importance[n]);
 * normalize != 0 divides channels >= normalize_col by d_normalizer[o] (or by the row
 * length when d_normalizer is NULL) where non-zero; then + bias, ReLU if requested.
 * Channel counts must be multiples of 4; all float pointers 16-byte aligned.
 * d_packed_filters (may be NULL): the filter bank packed by asr_pack_conv_filters
 * (asr_packed_conv_filters_size floats; hi/lo tf32 parts in the UMMA canonical layout).  When
 * given (out_channels <= 256) the contraction runs on the tensor cores (tcgen05.mma kind::tf32,
 * 3xTF32 split, fp32-level accuracy) and d_filters may be NULL; otherwise the fp32 FMA kernel runs. */
int64_t asr_packed_conv_filters_size(int kernel_size, int in_channels, int out_channels);
int asr_pack_conv_filters(const float* d_filters, int kernel_size, int in_channels, int out_channels,
                          float* d_packed, void* stream);
int asr_sparse_conv(const asr_conv_plan* plan, const float* d_filters, const float* d_packed_filters,
                    const float* d_inp_features,
                    int in_channels, int out_channels, const float* d_inp_importance,
                    const float* d_neighbors_importance, int importance_col, int normalize, int normalize_col,
                    const float* d_normalizer, const int64_t* d_neighbors_row_splits, const float* d_bias, int relu,
                    float* d_out_features, void* stream);
/* out[o] = sum_{n in row o} values[index ? index[n] : n]  (ops.reduce_subarrays_sum, common_torch.py:124-128) */
int asr_reduce_subarrays_sum(const float* d_values, const int32_t* d_index, const int64_t* d_row_splits,
                             int64_t num_rows, float* d_out, void* stream);
/* ops.invert_neighbors_list, net_definitions_torch.py:30-35 */
int asr_invert_neighbors_list(int64_t num_points, const int32_t* d_inp_neighbors_index,
                              const int64_t* d_inp_neighbors_row_splits, int64_t num_queries, int64_t num_entries,
                              const void* d_inp_attributes, int attribute_bytes, int32_t* d_neighbors_index,
                              int64_t* d_neighbors_row_splits, void* d_neighbors_attributes, void* stream);

/* ---------------------------------------------------------------- U-Net convolutions on split-half activations ("gx")
 * The fast path of model.unet (reference UNet5.unet, net_definitions_torch.py:535-638; every SparseConvBlock /
 * SparseConvTransitionBlock conv :123-387 = SpecialSparseConv.forward, models/common_torch.py:95-148, i.e. Open3D
 * `sparse_conv` + bias + ReLU): the same convolution as asr_sparse_conv for activations that stay on the device
 * between layers in a split-half format — a [V, C] tensor is stored as rows of `pitch` fp16 values holding hi =
 * fp16(x) at columns [hi, hi + C) and lo = fp16(x - hi) at [lo, lo + C) (4 bytes per element like fp32); buffers
 * have V + 1 rows, the last one all zero.  csrc/spconv_gx.cu describes the kernel (cp.async row gather, fp16 hi/lo
 * tcgen05 MMAs into TMEM, output-stationary dense slots + pair-major rare slots, fused and coalesced epilogue).
 * mode 0: within-grid (K = 55) and down (K = 9, inverted up table) tables; mode 1: up tables (one entry per row).
 * _plan_begin queues the counting kernels, _plan_finish (one host synchronisation, shared by all plans begun
 * before it) completes the plan and returns the number of rare entries (pair-buffer rows).  A plan may be built on
 * another stream than the one its convolutions run on (the caller orders the two with an event): from its first
 * convolution on, its buffers are released in the order of that convolution's stream. */
typedef struct asr_gx_plan asr_gx_plan;
/* d_row_map (may be NULL): the plan covers only the table rows d_row_map[0 .. num_out) (one rank's rows of a
 * sharded grid level, ascending); outputs / normalisers / residuals stay indexed by table row.  The array must
 * outlive the plan. */
int asr_gx_plan_begin(const int32_t* d_neighbors_index, const uint8_t* d_neighbors_kernel_index,
                      const int64_t* d_neighbors_row_splits, int64_t num_out, int64_t num_in, int64_t num_entries,
                      int kernel_size, int mode, const int32_t* d_row_map, void* stream, asr_gx_plan** out);
int asr_gx_plan_finish(asr_gx_plan* plan, void* stream, int64_t* num_rare);
void asr_gx_plan_destroy(asr_gx_plan* plan);
/* filters [K, Cin, Cout] fp32 -> fp16 hi/lo of W * 2^scale_exp for output columns [col0, col0 + ncols) in the
 * kernel's shared-memory image; Cin = 32 or a multiple of 64, ncols a multiple of 8, <= 256 */
int64_t asr_gx_packed_filters_bytes(int kernel_size, int in_channels, int ncols);
int asr_gx_pack_filters(const float* d_filters, int kernel_size, int in_channels, int out_channels, int col0, int ncols,
                        int scale_exp, void* d_packed, void* stream);
/* fp32 [num_rows, C] (row stride ldx) (* row_scale[row] if given) -> split-half view with out_rows rows; input row i
 * goes to output row d_rows[i] (d_rows NULL: i, out_rows = num_rows); also zeroes the view's zero row (row out_rows) */
int asr_gx_from_f32(const float* d_x, int64_t num_rows, int channels, int ldx, const float* d_row_scale,
                    const int32_t* d_rows, int64_t out_rows, void* d_out, int out_pitch, int out_hi, int out_lo,
                    void* stream);
int asr_gx_to_f32(const void* d_x, int64_t num_rows, int channels, int pitch, int hi, int lo, float* d_out, int ldo,
                  void* stream);
/* out = row_scale[row] * x, both split-half (the importance-weighted copy of SpecialSparseConv's conv1b input) */
int asr_gx_scale_rows(const void* d_x, int64_t num_rows, int channels, int pitch, int hi, int lo,
                      const float* d_row_scale, void* d_out, int out_pitch, int out_hi, int out_lo, void* stream);
/* out[o, 0:ncols] = act(sum_n (imp[idx_n]) x[idx_n] @ W[slot_n] (/ norm[o] where != 0) + bias) (+ residual[o]);
 * d_imp (may be NULL; ncols <= 128): importance per INPUT row, applied in fp32 in the epilogue; exactly one of
 * d_out_h2 (split-half view) / d_out_f32 (row stride out_f32_pitch floats) is given; d_pairbuf = scratch of
 * num_rare * roundup(ncols, 16) floats (mode 0 plans with rare entries) */
int asr_gx_conv(const asr_gx_plan* plan, const void* d_x, int in_channels, int x_pitch, int x_hi, int x_lo,
                const void* d_packed, int ncols, int scale_exp, const float* d_bias, int relu, const float* d_norm,
                const float* d_imp, const void* d_res, int res_pitch, int res_hi, int res_lo, void* d_out_h2, int out_pitch, int out_hi,
                int out_lo, float* d_out_f32, int out_f32_pitch, float* d_pairbuf, void* stream);
/* 1 if a conversion to the split-half format saturated (|x| > 65504) since the last call; synchronises */
int asr_gx_overflow(void* stream, int* flag);
/* Development aid (option "gx_trace" = 1): 16 cycle counters per CTA (first `ctas` <= 256 CTAs) of the most recent
 * asr_gx_conv kernel launch — how long one thread of every warp role waited on each of its barriers; the slots
 * are listed in csrc/spconv_gx.cu.  Host array of 16 * ctas unsigned.  No reference counterpart. */
int asr_gx_trace(void* stream, int ctas, unsigned* counters);

/* ---------------------------------------------------------------- multi-GPU sharding helpers
 * The reference is single-process; these serve the sharded form of the path (one process per GPU, geometry
 * replicated, every grid level's rows owned by Z-curve region, csrc/shard.cu).  All asynchronous on `stream`.
 * _positions: Z-curve position (depth 21) of every location code; _owner: owner rank of every voxel = number of
 * thresholds <= its position (thresholds = region boundaries, ascending, world - 1 of them);
 * _need_mask: ORs into d_mask[row of the input level] the ranks (bit r) whose output rows read that row through the
 * neighbour table, for rows owned by `rank`; _push: copies the listed (owned) rows — the bytes [seg_off[i],
 * seg_off[i] + seg_len) of each row — into the same buffer on every rank flagged in d_mask (NULL: all ranks) through
 * the peer-mapped base pointers of a symmetric allocation (buffer = base + offset on every rank). */
int asr_shard_positions(const uint64_t* d_keys, int64_t num_voxels, uint64_t* d_pos, void* stream);
int asr_shard_owner(const uint64_t* d_keys, int64_t num_voxels, const uint64_t* d_thresholds, int num_thresholds,
                    uint8_t* d_owner, void* stream);
int asr_shard_need_mask(const int64_t* d_row_splits, const int32_t* d_index, int64_t num_out, const uint8_t* d_owner_out,
                        const uint8_t* d_owner_in, int rank, uint32_t* d_mask, void* stream);
int asr_shard_push(void* const* peer_base, int world, int rank, int64_t offset, int64_t pitch, const int* seg_off, int nseg,
                   int seg_len, const int32_t* d_rows, int64_t num_rows, const uint32_t* d_mask, void* stream);

/* ---------------------------------------------------------------- decoder MLP
 * replaces UNet5.decode / decode_with_gradient, net_definitions_torch.py:655-686.
 * weights in torch.nn.Linear layout: w1 [32,35], b1 [32], w2 [32,32], b2 [32], w3 [2,32].
 * d_shifts may be NULL (zeros, as asr.cpp:324); d_signed_scale (may be NULL)
 * multiplies channel 0 per voxel (asr.cpp:334-336); d_grad (may be NULL) receives
 * d value[:,0] / d shift. */
int asr_decode(const float* d_shifts, const float* d_code, int64_t num_voxels, const float* d_w1, const float* d_b1,
               const float* d_w2, const float* d_b2, const float* d_w3, const float* d_signed_scale,
               float* d_values, float* d_grad, void* stream);

/* ---------------------------------------------------------------- dense layer on tensor cores
 * out[rows, out_features] = act(a[rows, in_features] @ w[in_features, out_features] + bias) on
 * tcgen05.mma kind::tf32 with a 3xTF32 split (fp32-level accuracy), accumulators in TMEM.  The
 * weight matrix ([in, out] row-major, i.e. torch.nn.Linear.weight.T) is packed once with
 * asr_pack_weights into asr_packed_weights_size() floats.  out_features <= 256.  d_bias may be
 * NULL; relu applies only with a bias.  Used for the decoder MLP layers
 * (net_definitions_torch.py:503-511,655-666). */
int64_t asr_packed_weights_size(int in_features, int out_features);
int asr_pack_weights(const float* d_w, int in_features, int out_features, float* d_packed, void* stream);
int asr_dense_tf32x3(const float* d_a, int64_t rows, int in_features, int lda, const float* d_packed_w,
                     int out_features, const float* d_bias, int relu, float* d_out, int ldd, void* stream);

/* ---------------------------------------------------------------- dual contouring (vertices)
 * replaces the vertex passes of asr::CreateTriangleMesh, cpp/lib/contouring.h:25-30 /
 * contouring.cpp:66-199.  values [V,2] (signed, unsigned), dual indices [D,8],
 * node positions [V,3].  _count fills d_flag[D] (u8) and d_offset[D+1] (i64) scratch
 * owned by the caller and returns the number of vertices; _fill writes
 * vertices [M,3] and (optionally) the dual index of every vertex [M]. */
int asr_contour_count(const float* d_values, const int64_t* d_dual_indices, int64_t num_duals,
                      float unsigned_threshold, uint8_t* d_flag, int64_t* d_offset, int64_t* num_vertices,
                      void* stream);
int asr_contour_fill(const float* d_values, const int64_t* d_dual_indices, int64_t num_duals,
                     float unsigned_threshold, const float* d_node_positions, const uint8_t* d_flag,
                     const int64_t* d_offset, float* d_vertices, int64_t* d_vertex_dual, void* stream);

/* ---------------------------------------------------------------- dual contouring (triangles)
 * replaces the polygon passes of asr::CreateTriangleMesh, cpp/lib/contouring.cpp:202-459: one
 * polygon per sign-changing primal edge, built from the crossing dual cells around it (1 triangle,
 * 2 triangles split along the shorter diagonal, or a fan around an extra centre vertex).
 * d_vertex_dual [M] is the output of asr_contour_fill.  _create returns the number of triangles T
 * and of extra (fan centre) vertices X; _fill takes d_vertices [(M + X), 3] whose first M rows
 * hold the dual-cell vertices, appends the X centre vertices and writes d_triangles [T, 3]
 * (int32, emission order of the reference: vertex-major, edge-minor; the rotation of each
 * triangle / fan is unspecified in the reference, see contour_tri.cu).  The value / dual /
 * vertex_dual buffers passed to _create must stay valid until _fill. */
int asr_contour_triangles_create(const float* d_values, const int64_t* d_dual_indices, int64_t num_duals,
                                 float unsigned_threshold, const int64_t* d_vertex_dual, int64_t num_vertices,
                                 int64_t num_nodes, void* stream, void** handle, int64_t* num_triangles,
                                 int64_t* num_extra_vertices);
int asr_contour_triangles_fill(void* handle, float* d_vertices, int32_t* d_triangles, void* stream);
void asr_contour_triangles_destroy(void* handle);

/* ---------------------------------------------------------------- point pre-processing (row f-2)
 * replaces asr::KDTree (cpp/lib/nsearch.cpp:22-105; python KDTree, module.cpp:455-489).
 * _k_radius: sqrt of the largest of the k smallest squared distances, the point itself included
 * (ComputeKRadius :30-52); _inlier: 1 unless at least `outlier_threshold` of the k nearest points
 * have a radius < radius_fraction * own radius (ComputeInlier :54-85); counts: number of points
 * with |p - p_i|^2 < r_i^2 (ComputeRadiusNeighbors :87-105).  k <= 64.  The handle is destroyed
 * with asr_radius_search_destroy. */
int asr_kdtree_create(const float* d_points, int64_t num_points, void* stream, asr_search** out);
int asr_kdtree_k_radius(asr_search* tree, int k, float* d_out, void* stream);
int asr_kdtree_inlier(asr_search* tree, const float* d_radii, float radius_fraction, int k, int outlier_threshold,
                      uint8_t* d_out, void* stream);
int asr_radius_neighbor_counts(const float* d_points, int64_t num_points, const float* d_radii, int32_t* d_out,
                               void* stream);

/* ---------------------------------------------------------------- mesh post-processing
 * replaces asr::ConnectedComponents (cpp/lib/postprocess.cpp:81-141), the core of
 * RemoveConnectedComponents (:143-176; python remove_connected_components, module.cpp:348).
 * d_label [V]: smallest vertex index of the vertex' component (orders components like the
 * reference's first-visit numbering); d_size [V] (optional): component size at the label's
 * index, 0 elsewhere.  Out-of-range triangle indices -> status 2. */
int asr_mesh_components(const int32_t* d_triangles, int64_t num_triangles, int64_t num_vertices, int64_t* d_label,
                        int64_t* d_size, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif
#endif /* ASR_B200_H */
