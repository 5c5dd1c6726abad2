"""Tensor-level host wrappers over the C ABI (include/asr_b200.h).

PyTorch is only the plumbing here: it owns device memory and the CUDA stream;
every computation is one of the library's hand-written kernels.  All tensors must
live on the current CUDA device; a CPU tensor raises (there is no fallback).
"""
import ctypes as C
from collections import namedtuple

import numpy as np
import torch

from ._lib import check, lib

_i64 = C.c_int64

# bench.py sets this to a list to collect the shape of every sparse conv of one
# step (for the algorithmic byte / flop accounting); None = off.
ACCOUNT = None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _cuda(t, dtype, name):
    if not isinstance(t, torch.Tensor):
        raise ValueError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("asr_b200: %s is on %s — the hot path runs on CUDA only (no CPU fallback)" %
                           (name, t.device))
    if t.dtype != dtype:
        t = t.to(dtype)
    t = t.detach().contiguous()
    if t.data_ptr() % 16:
        t = t.clone()
    return t


def _opt(t, dtype, name):
    if t is None or t.numel() == 0:
        return None
    return _cuda(t, dtype, name)


# ------------------------------------------------------------------------------------ octree / grids

GRID_FIELDS = ("voxel_keys", "voxel_centers", "voxel_sizes", "neighbors_index", "neighbors_kernel_index",
               "neighbors_row_splits", "up_neighbors_index", "up_neighbors_kernel_index", "up_neighbors_row_splits")


class Octree:
    """Opaque octree handle (mirror of the python module's `Octree`, module.cpp:282),
    built on the GPU by asr_octree_create (reference CreateOctreeFromPoints, octree.cpp:230)."""

    def __init__(self, points, radii, bb_min, bb_max, radius_scale=1.0, grow_steps=0, max_depth=21):
        points = _cuda(points, torch.float32, "points")
        radii = _cuda(radii, torch.float32, "radii")
        if points.ndim != 2 or points.shape[1] != 3:
            raise ValueError("points must have shape [N,3]")
        if radii.ndim != 1 or radii.shape[0] != points.shape[0]:
            raise ValueError("radii must have shape [N]")
        bb_min = np.ascontiguousarray(np.asarray(bb_min, dtype=np.float32).reshape(3))
        bb_max = np.ascontiguousarray(np.asarray(bb_max, dtype=np.float32).reshape(3))
        self.device = points.device
        self._h = C.c_void_p(0)
        check(lib().asr_octree_create(_ptr(points), _ptr(radii), points.shape[0], bb_min.ctypes.data,
                                      bb_max.ctypes.data, float(radius_scale), int(grow_steps), int(max_depth),
                                      _stream(), C.byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                lib().asr_octree_destroy(h)
            except Exception:
                pass
            self._h = None

    @property
    def num_leaves(self):
        return int(lib().asr_octree_num_leaves(self._h))

    @property
    def num_nodes(self):
        return int(lib().asr_octree_num_nodes(self._h))

    @property
    def balance_rounds(self):
        return int(lib().asr_octree_balance_rounds(self._h))

    def leaves(self):
        out = torch.empty(self.num_leaves, dtype=torch.int64, device=self.device)
        check(lib().asr_octree_get_leaves(self._h, _ptr(out), _stream()))
        return out  # bit pattern of the uint64 location codes

    def frame(self):
        vs = np.empty(22, np.float32)
        ivs = np.empty(22, np.float32)
        off = np.empty(3, np.int32)
        check(lib().asr_octree_get_frame(self._h, vs.ctypes.data, ivs.ctypes.data, off.ctypes.data))
        return vs, ivs, off

    def search_frame(self):
        """(origin x, y, z, finest cell size) of the octree's 2^21 grid, for multi_radius_search."""
        vs, _, off = self.frame()
        h = np.float32(vs[21])
        return np.array([-np.float32(off[0]) * h, -np.float32(off[1]) * h, -np.float32(off[2]) * h, h], np.float32)

    def grids(self, num_levels, voxel_info_all_levels=False):
        """asr_grids_* (reference CreateGridsFromOctree, grid.cpp:245).  Returns a
        list (finest first) of dicts of CUDA tensors with the key set of
        pyCreateGridsFromOctree (module.cpp:163-228); voxel_keys is int64 holding
        the uint64 bit pattern."""
        L = lib()
        check(L.asr_grids_build(self._h, int(num_levels), int(bool(voxel_info_all_levels)), _stream()))
        dev = self.device
        res = []
        for lev in range(num_levels):
            V, E = _i64(0), _i64(0)
            check(L.asr_grids_level_size(self._h, lev, C.byref(V), C.byref(E)))
            V, E = V.value, E.value
            info = lev == 0 or voxel_info_all_levels
            up = lev < num_levels - 1
            g = {}
            if info:
                g["voxel_keys"] = torch.empty(V, dtype=torch.int64, device=dev)
                g["voxel_centers"] = torch.empty((V, 3), dtype=torch.float32, device=dev)
                g["voxel_sizes"] = torch.empty(V, dtype=torch.float32, device=dev)
            g["neighbors_index"] = torch.empty(E, dtype=torch.int32, device=dev)
            g["neighbors_kernel_index"] = torch.empty(E, dtype=torch.uint8, device=dev)
            g["neighbors_row_splits"] = torch.empty(V + 1, dtype=torch.int64, device=dev)
            if up:
                g["up_neighbors_index"] = torch.empty(V, dtype=torch.int32, device=dev)
                g["up_neighbors_kernel_index"] = torch.empty(V, dtype=torch.uint8, device=dev)
                g["up_neighbors_row_splits"] = torch.empty(V + 1, dtype=torch.int64, device=dev)
            check(L.asr_grids_get(self._h, lev, *[_ptr(g.get(k)) for k in GRID_FIELDS], _stream()))
            # the reference omits empty vectors from the dict (module.cpp:174-223)
            res.append({k: v for k, v in g.items() if v.numel() > 0})
        return res

    def dual_vertex_indices(self):
        """[num_duals, 8] int64 leaf indices (reference CreateDualVertexIndices, grid.cpp:450)."""
        out = self.dual_vertex_indices_finish()
        check(lib().asr_duals_check(self._h))
        return out

    # asynchronous form for the pipeline: begin (counting pass queued, no host synchronisation) ... finish (waits for
    # the count, queues the fill) ... dual_check() before the result is trusted.  Both may run on a side stream.
    def dual_vertex_indices_begin(self):
        check(lib().asr_duals_begin(self._h, _stream()))

    def dual_vertex_indices_finish(self):
        n = _i64(0)
        check(lib().asr_duals_count(self._h, C.byref(n), _stream()))
        out = torch.empty((n.value, 8), dtype=torch.int64, device=self.device)
        check(lib().asr_duals_fill(self._h, _ptr(out), _stream()))
        return out

    def dual_check(self):
        check(lib().asr_duals_check(self._h))


# ------------------------------------------------------------------------------------ aggregation search


def multi_radius_search(points, queries, radii, frame=None):
    """(index int32[P], squared distance float32[P], row_splits int64[Q+1]);
    d2 < r^2, ascending by d2 (reference nsearch.cpp:130-146).  `frame` =
    (origin xyz, finest cell size) optionally aligns the internal binning grid
    with the octree (Octree.search_frame()); it never changes the result."""
    points = _cuda(points, torch.float32, "points")
    queries = _cuda(queries, torch.float32, "queries")
    radii = _cuda(radii, torch.float32, "radii")
    if points.ndim != 2 or points.shape[1] != 3 or queries.ndim != 2 or queries.shape[1] != 3:
        raise ValueError("points and queries must have shape [N,3]")
    if radii.shape != (queries.shape[0],):
        raise ValueError("radii must have shape [num_queries]")
    h, n = C.c_void_p(0), _i64(0)
    L = lib()
    fr = None
    if frame is not None:
        fr = np.ascontiguousarray(np.asarray(frame, dtype=np.float32).reshape(4))
    check(L.asr_radius_search_create(_ptr(points), points.shape[0], _ptr(queries), _ptr(radii), queries.shape[0],
                                     fr.ctypes.data if fr is not None else None, _stream(), C.byref(h), C.byref(n)))
    try:
        dev = points.device
        idx = torch.empty(n.value, dtype=torch.int32, device=dev)
        dist = torch.empty(n.value, dtype=torch.float32, device=dev)
        rs = torch.empty(queries.shape[0] + 1, dtype=torch.int64, device=dev)
        check(L.asr_radius_search_fill(h, _ptr(idx), _ptr(dist), _ptr(rs), _stream()))
    finally:
        L.asr_radius_search_destroy(h)
    return idx, dist, rs


def scale_compatibility(voxel_sizes, point_radii, neighbors_index, neighbors_row_splits):
    voxel_sizes = _cuda(voxel_sizes, torch.float32, "voxel_sizes")
    point_radii = _cuda(point_radii, torch.float32, "point_radii")
    idx = _cuda(neighbors_index, torch.int32, "neighbors_index")
    rs = _cuda(neighbors_row_splits, torch.int64, "neighbors_row_splits")
    out = torch.empty(idx.shape[0], dtype=torch.float32, device=idx.device)
    check(lib().asr_scale_compatibility(_ptr(voxel_sizes), _ptr(point_radii), _ptr(idx), _ptr(rs),
                                        voxel_sizes.shape[0], _ptr(out), _stream()))
    return out


def aggregation_importance(scale_compat, dist):
    scale_compat = _cuda(scale_compat, torch.float32, "scale_compat")
    dist = _cuda(dist, torch.float32, "dist")
    out = torch.empty_like(dist)
    check(lib().asr_aggregation_importance(_ptr(scale_compat), _ptr(dist), dist.shape[0], _ptr(out), _stream()))
    return out


# ------------------------------------------------------------------------------------ convolutions


def _out_buffer(out, rows, cols, device):
    if out is None:
        return torch.empty((rows, cols), dtype=torch.float32, device=device)
    if (not out.is_cuda or out.dtype != torch.float32 or tuple(out.shape) != (rows, cols) or not out.is_contiguous()
            or out.data_ptr() % 16):
        raise ValueError("out must be a contiguous 16-byte aligned CUDA float32 tensor of shape [%d, %d]" % (rows, cols))
    return out


def pair_importance_for_unet(importance, num_voxels):
    """Identity on one GPU; the sharded namespace (shard.ShardedOps) assembles the
    first `num_voxels` entries of the global pair list here."""
    return importance


def continuous_conv(filters, out_positions, extents, offset, inp_positions, inp_features, inp_importance,
                    neighbors_index, neighbors_importance, neighbors_row_splits, normalize=True, bias=None,
                    relu=False, out=None):
    filters = _cuda(filters, torch.float32, "filters")
    if filters.ndim != 5 or not (filters.shape[0] == filters.shape[1] == filters.shape[2]):
        raise ValueError("filters must have shape [S,S,S,Cin,Cout]")
    S, _, _, Cin, Cout = filters.shape
    out_positions = _cuda(out_positions, torch.float32, "out_positions")
    extents = _cuda(extents, torch.float32, "extents").reshape(-1)
    V = out_positions.shape[0]
    if extents.numel() not in (1, V):
        raise ValueError("extents must have 1 or num_out elements")
    inp_positions = _cuda(inp_positions, torch.float32, "inp_positions")
    inp_features = _cuda(inp_features, torch.float32, "inp_features")
    if inp_features.shape[1] != Cin:
        raise ValueError("inp_features channel count does not match the filter")
    idx = _cuda(neighbors_index, torch.int32, "neighbors_index")
    rs = _cuda(neighbors_row_splits, torch.int64, "neighbors_row_splits")
    if rs.shape[0] != V + 1:
        raise ValueError("neighbors_row_splits must have num_out+1 elements")
    out = _out_buffer(out, V, Cout, filters.device)
    check(lib().asr_continuous_conv(
        _ptr(filters), _ptr(out_positions), _ptr(extents), 1 if extents.numel() == V and V != 1 else 0,
        _ptr(_opt(offset, torch.float32, "offset")), _ptr(inp_positions), _ptr(inp_features),
        _ptr(_opt(inp_importance, torch.float32, "inp_importance")), _ptr(idx),
        _ptr(_opt(neighbors_importance, torch.float32, "neighbors_importance")), _ptr(rs), V, S, Cin, Cout,
        int(bool(normalize)), _ptr(_opt(bias, torch.float32, "bias")), int(bool(relu)), _ptr(out), _stream()))
    return out


def cat(tensors, dim=-1):
    """Channel concatenation of the U-Net's skip connections (the sharded namespace keeps its
    row-validity bookkeeping here)."""
    return torch.cat(tensors, dim)


def add(x, y):
    return x + y


class ConvPlan:
    """Slot-sorted form of one neighbour table; build once, reuse for every conv on it."""

    def __init__(self, neighbors_index, neighbors_kernel_index, neighbors_row_splits, kernel_size):
        self.idx = _cuda(neighbors_index, torch.int32, "neighbors_index")
        self.slot = _cuda(neighbors_kernel_index, torch.uint8, "neighbors_kernel_index")
        self.row_splits = _cuda(neighbors_row_splits, torch.int64, "neighbors_row_splits")
        if self.idx.shape != self.slot.shape:
            raise ValueError("neighbors_index and neighbors_kernel_index must have the same length")
        self.num_out = self.row_splits.shape[0] - 1
        self.kernel_size = int(kernel_size)
        self._h = C.c_void_p(0)
        check(lib().asr_conv_plan_create(_ptr(self.idx), _ptr(self.slot), _ptr(self.row_splits), self.num_out,
                                         self.idx.shape[0], self.kernel_size, _stream(), C.byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                lib().asr_conv_plan_destroy(h)
            except Exception:
                pass
            self._h = None


# Backend of the U-Net's sparse convolutions:
#   "gx"     (default) split-half activations between layers, TMA row gather + fp16 hi/lo tcgen05 MMAs,
#            output-stationary dense slots (csrc/spconv_gx.cu; model.UNet.unet only)
#   "tensor" pair-major tcgen05 3xTF32 kernel on fp32 activations (round 1; also what the generic
#            asr_sparse_conv / open3d::sparse_conv entry points run — the gx setting falls back to it there)
#   "fp32"   the FMA tile kernel
SPARSE_CONV_BACKEND = "gx"


class PackedFilters:
    """A [K, Cin, Cout] filter bank packed for the tensor-core tile kernel."""

    def __init__(self, filters):
        filters = _cuda(filters, torch.float32, "filters")
        self.shape = tuple(filters.shape)
        K, Cin, Cout = self.shape
        if Cout > 256:
            raise ValueError("tensor-core sparse conv needs out_channels <= 256")
        self.data = torch.empty(int(lib().asr_packed_conv_filters_size(K, Cin, Cout)), dtype=torch.float32,
                                device=filters.device)
        check(lib().asr_pack_conv_filters(_ptr(filters), K, Cin, Cout, _ptr(self.data), _stream()))


def sparse_conv(plan, filters, inp_features, inp_importance=None, neighbors_importance=None, importance_col=0,
                normalize=False, normalize_col=0, normalizer=None, bias=None, relu=False, backend=None, out=None):
    """out[o] = sum_n imp_n * x[idx_n] @ filters[slot_n] (+ normalise, bias, ReLU).
    `filters` is a [K, Cin, Cout] tensor or a PackedFilters (pre-packed, tensor cores)."""
    backend = backend or SPARSE_CONV_BACKEND
    if backend == "gx":
        backend = "tensor"
    packed = None
    if isinstance(filters, PackedFilters):
        packed, filters = filters, None
        K, Cin, Cout = packed.shape
    else:
        filters = _cuda(filters, torch.float32, "filters")
        K, Cin, Cout = filters.shape
        if backend == "tensor" and Cout <= 256:
            packed = PackedFilters(filters)
    x = _cuda(inp_features, torch.float32, "inp_features")
    if K != plan.kernel_size:
        raise ValueError("filters.shape[0] does not match the plan's kernel size")
    if x.shape[1] != Cin:
        raise ValueError("inp_features channel count does not match the filter")
    if ACCOUNT is not None:
        ACCOUNT.append({"V_in": x.shape[0], "V_out": plan.num_out, "E": plan.idx.shape[0], "K": K, "Cin": Cin,
                        "Cout": Cout, "importance": inp_importance is not None or neighbors_importance is not None})
    out = _out_buffer(out, plan.num_out, Cout, x.device)
    check(lib().asr_sparse_conv(plan._h, _ptr(filters), _ptr(packed.data) if packed is not None else None, _ptr(x),
                                Cin, Cout,
                                _ptr(_opt(inp_importance, torch.float32, "inp_importance")),
                                _ptr(_opt(neighbors_importance, torch.float32, "neighbors_importance")),
                                int(importance_col), int(bool(normalize)), int(normalize_col),
                                _ptr(_opt(normalizer, torch.float32, "normalizer")), _ptr(plan.row_splits),
                                _ptr(_opt(bias, torch.float32, "bias")), int(bool(relu)), _ptr(out), _stream()))
    return out


def reduce_subarrays_sum(values, row_splits, index=None):
    values = _cuda(values, torch.float32, "values")
    rs = _cuda(row_splits, torch.int64, "row_splits")
    idx = _opt(index, torch.int32, "index")
    out = torch.empty(rs.shape[0] - 1, dtype=torch.float32, device=values.device)
    check(lib().asr_reduce_subarrays_sum(_ptr(values), _ptr(idx), _ptr(rs), rs.shape[0] - 1, _ptr(out), _stream()))
    return out


InvertNeighborsListResult = namedtuple("InvertNeighborsListResult",
                                       ["neighbors_index", "neighbors_row_splits", "neighbors_attributes"])


def invert_neighbors_list(num_points, inp_neighbors_index, inp_neighbors_row_splits, inp_neighbors_attributes):
    idx = _cuda(inp_neighbors_index, torch.int32, "inp_neighbors_index")
    rs = _cuda(inp_neighbors_row_splits, torch.int64, "inp_neighbors_row_splits")
    attrs = inp_neighbors_attributes
    has_attr = attrs is not None and attrs.numel() > 0
    if has_attr:
        attrs = _cuda(attrs, attrs.dtype, "inp_neighbors_attributes")
        if attrs.shape[0] != idx.shape[0]:
            raise ValueError("attributes must have one entry per neighbour")
    dev = idx.device
    out_idx = torch.empty_like(idx)
    out_rs = torch.empty(num_points + 1, dtype=torch.int64, device=dev)
    out_attr = torch.empty_like(attrs) if has_attr else torch.empty(0, dtype=torch.float32, device=dev)
    abytes = attrs.element_size() * int(np.prod(attrs.shape[1:])) if has_attr else 0
    check(lib().asr_invert_neighbors_list(int(num_points), _ptr(idx), _ptr(rs), rs.shape[0] - 1, idx.shape[0],
                                          _ptr(attrs if has_attr else None), abytes, _ptr(out_idx), _ptr(out_rs),
                                          _ptr(out_attr if has_attr else None), _stream()))
    return InvertNeighborsListResult(out_idx, out_rs, out_attr)


# ------------------------------------------------------------------------------------ decode / contouring


def decode(shifts, code, w1, b1, w2, b2, w3, signed_scale=None, with_gradient=False):
    code = _cuda(code, torch.float32, "code")
    if code.ndim != 2 or code.shape[1] != 32:
        raise ValueError("code must have shape [V,32]")
    V = code.shape[0]
    shifts = _opt(shifts, torch.float32, "shifts")
    if shifts is not None and shifts.shape != (V, 3):
        raise ValueError("shifts must have shape [V,3]")
    ws = [_cuda(w, torch.float32, "decoder weight") for w in (w1, b1, w2, b2, w3)]
    if ws[0].shape != (32, 35) or ws[2].shape != (32, 32) or ws[4].shape != (2, 32):
        raise ValueError("decoder weights must have shapes [32,35], [32,32], [2,32]")
    values = torch.empty((V, 2), dtype=torch.float32, device=code.device)
    grad = torch.empty((V, 3), dtype=torch.float32, device=code.device) if with_gradient else None
    check(lib().asr_decode(_ptr(shifts), _ptr(code), V, *[_ptr(w) for w in ws],
                           _ptr(_opt(signed_scale, torch.float32, "signed_scale")), _ptr(values), _ptr(grad),
                           _stream()))
    return (values, grad) if with_gradient else values


def pack_weights(w):
    """Packs a [in, out] fp32 matrix for dense_tf32x3 (hi/lo tf32 parts, UMMA canonical layout)."""
    w = _cuda(w, torch.float32, "w")
    K, N = w.shape
    out = torch.empty(int(lib().asr_packed_weights_size(K, N)), dtype=torch.float32, device=w.device)
    check(lib().asr_pack_weights(_ptr(w), K, N, _ptr(out), _stream()))
    return out, K, N


def dense_tf32x3(a, packed, bias=None, relu=False):
    """act(a @ w + bias) on the tensor cores (tcgen05 kind::tf32, 3xTF32 split)."""
    wp, K, N = packed
    a = _cuda(a, torch.float32, "a")
    if a.ndim != 2 or a.shape[1] != K:
        raise ValueError("a must have shape [rows, %d]" % K)
    out = torch.empty((a.shape[0], N), dtype=torch.float32, device=a.device)
    check(lib().asr_dense_tf32x3(_ptr(a), a.shape[0], K, K, _ptr(wp), N, _ptr(_opt(bias, torch.float32, "bias")),
                                 int(bool(relu)), _ptr(out), N, _stream()))
    return out


def contour_vertices(values, dual_indices, node_positions, unsigned_threshold=1.0):
    """Vertex part of CreateTriangleMesh (contouring.cpp:66-199).  Returns
    (vertices f32[M,3], dual index of each vertex i64[M])."""
    values = _cuda(values, torch.float32, "values")
    duals = _cuda(dual_indices, torch.int64, "dual_indices")
    pos = _cuda(node_positions, torch.float32, "node_positions")
    if values.ndim != 2 or values.shape[1] != 2:
        raise RuntimeError("values vector size is not a multiple of 2.")
    if duals.ndim != 2 or duals.shape[1] != 8:
        raise RuntimeError("dual_indices vector size is not a multiple of 8.")
    if pos.ndim != 2 or pos.shape[1] != 3:
        raise RuntimeError("node_positions vector size is not a multiple of 3.")
    D = duals.shape[0]
    dev = values.device
    flag = torch.empty(D, dtype=torch.uint8, device=dev)
    offset = torch.empty(D + 1, dtype=torch.int64, device=dev)
    n = _i64(0)
    L = lib()
    check(L.asr_contour_count(_ptr(values), _ptr(duals), D, float(unsigned_threshold), _ptr(flag), _ptr(offset),
                              C.byref(n), _stream()))
    verts = torch.empty((n.value, 3), dtype=torch.float32, device=dev)
    vdual = torch.empty(n.value, dtype=torch.int64, device=dev)
    check(L.asr_contour_fill(_ptr(values), _ptr(duals), D, float(unsigned_threshold), _ptr(pos), _ptr(flag),
                             _ptr(offset), _ptr(verts), _ptr(vdual), _stream()))
    return verts, vdual


def contour_mesh(values, dual_indices, node_positions, unsigned_threshold=1.0):
    """CreateTriangleMesh (contouring.cpp:29-460) on the GPU.  Returns (vertices f32[M + X, 3],
    triangles i32[T, 3], dual index of the first M vertices i64[M]); X = fan-centre vertices."""
    verts, vdual = contour_vertices(values, dual_indices, node_positions, unsigned_threshold)
    values = _cuda(values, torch.float32, "values")
    duals = _cuda(dual_indices, torch.int64, "dual_indices")
    M = verts.shape[0]
    L = lib()
    h, T, X = C.c_void_p(0), _i64(0), _i64(0)
    check(L.asr_contour_triangles_create(_ptr(values), _ptr(duals), duals.shape[0], float(unsigned_threshold),
                                         _ptr(vdual), M, values.shape[0], _stream(), C.byref(h), C.byref(T),
                                         C.byref(X)))
    try:
        allv = torch.empty((M + X.value, 3), dtype=torch.float32, device=verts.device)
        allv[:M] = verts
        tris = torch.empty((T.value, 3), dtype=torch.int32, device=verts.device)
        check(L.asr_contour_triangles_fill(h, _ptr(allv), _ptr(tris), _stream()))
    finally:
        L.asr_contour_triangles_destroy(h)
    return allv, tris, vdual


def remove_connected_components(vertices, triangles, keep_n_largest_components, minimum_component_size=3):
    """RemoveConnectedComponents (postprocess.cpp:143-201): keep the `keep_n` largest vertex
    components (ties: the later component first, std::greater on (size, id)) that have at least
    `minimum_component_size` vertices; vertices keep their order, triangles are re-indexed.
    The labelling runs in asr_mesh_components; the selection below is index plumbing."""
    vertices = _cuda(vertices, torch.float32, "vertices")
    tris = _cuda(triangles, torch.int32, "triangles")
    if vertices.ndim != 2 or vertices.shape[1] != 3:
        raise ValueError("vertices must have shape [N,3]")
    if tris.ndim != 2 or tris.shape[1] != 3:
        raise ValueError("triangles must have shape [N,3]")
    V = vertices.shape[0]
    if V == 0:
        return vertices, tris
    label = torch.empty(V, dtype=torch.int64, device=vertices.device)
    size = torch.empty(V, dtype=torch.int64, device=vertices.device)
    check(lib().asr_mesh_components(_ptr(tris), tris.shape[0], V, _ptr(label), _ptr(size), _stream()))
    roots = torch.nonzero(size > 0).reshape(-1)  # ascending = the reference's component ids
    order = torch.argsort(size[roots] * (V + 1) + roots, descending=True)
    keep_n = int(min(int(keep_n_largest_components), roots.numel())) if keep_n_largest_components > 0 else 0
    chosen = roots[order[:keep_n]]
    chosen = chosen[size[chosen] >= int(minimum_component_size)]
    keep_root = torch.zeros(V, dtype=torch.bool, device=vertices.device)
    keep_root[chosen] = True
    mask = keep_root[label]
    new_index = torch.cumsum(mask, 0, dtype=torch.int64) - 1
    tl = tris.long()
    tmask = mask[tl[:, 0]] & mask[tl[:, 1]] & mask[tl[:, 2]]
    return vertices[mask].contiguous(), new_index[tl[tmask]].to(torch.int32).contiguous()


# ------------------------------------------------------------------------------------ point pre-processing (f-2)


class KDTree:
    """Nearest-neighbour queries of a cloud among itself (reference KDTree, nsearch.cpp:22-105)."""

    def __init__(self, points):
        self.points = _cuda(points, torch.float32, "points")
        if self.points.ndim != 2 or self.points.shape[1] != 3:
            raise ValueError("points must have shape [N,3]")
        self._h = C.c_void_p(0)
        check(lib().asr_kdtree_create(_ptr(self.points), self.points.shape[0], _stream(), C.byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                lib().asr_radius_search_destroy(h)
            except Exception:
                pass
            self._h = None

    def compute_k_radius(self, k):
        out = torch.empty(self.points.shape[0], dtype=torch.float32, device=self.points.device)
        check(lib().asr_kdtree_k_radius(self._h, int(k), _ptr(out), _stream()))
        return out

    def compute_inlier(self, radii, radius_fraction=0.5, k=24, outlier_threshold=1):
        radii = _cuda(radii, torch.float32, "radii")
        if radii.shape != (self.points.shape[0],):
            raise ValueError("radii must have shape [N]")
        out = torch.empty(self.points.shape[0], dtype=torch.uint8, device=self.points.device)
        check(lib().asr_kdtree_inlier(self._h, _ptr(radii), float(radius_fraction), int(k), int(outlier_threshold),
                                      _ptr(out), _stream()))
        return out.bool()

    def compute_radius_neighbors(self, radii):
        radii = _cuda(radii, torch.float32, "radii")
        if radii.shape != (self.points.shape[0],):
            raise ValueError("radii must have shape [N]")
        out = torch.empty(self.points.shape[0], dtype=torch.int32, device=self.points.device)
        check(lib().asr_radius_neighbor_counts(_ptr(self.points), self.points.shape[0], _ptr(radii), _ptr(out),
                                               _stream()))
        return out


def density_inlier(radius_neighbors, density_percentile_threshold):
    """Points whose radius-neighbour count exceeds the count at the percentile rank
    (ComputeInlierFromDensity, preprocess.cpp:38-62).  NOTE: the reference compares the
    *partially sorted* count array with the threshold (:50-61), so which points it keeps depends on
    std::partial_sort's unspecified remainder order; this keeps the documented intent — the same
    NUMBER of points, selected by their own count."""
    n = radius_neighbors.shape[0]
    middle = int((float(density_percentile_threshold) / 100.0) * n)
    middle = min(n, max(1, middle))
    threshold = torch.kthvalue(radius_neighbors.to(torch.int64), middle).values
    return radius_neighbors > threshold
