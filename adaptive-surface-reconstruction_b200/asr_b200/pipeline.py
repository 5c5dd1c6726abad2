"""The hot path end to end on one GPU: (points, normals, radii) -> octree -> grid
hierarchy + neighbour tables -> dual cells -> aggregation neighbours ->
aggregate / unet / decode -> dual-contouring vertices.

Python re-statement of the *sequencing* in asr::ReconstructSurface (reference
cpp/lib/asr.cpp:143-342); every stage is a kernel of libasr_b200.so.  The input
dict has exactly the keys the reference builds (asr.cpp:159-312).
"""
import time

import numpy as np
import torch

from . import ops


class StageTimer:
    """CUDA-event timer per pipeline stage (enabled on demand; adds syncs)."""

    def __init__(self, enabled=False):
        self.enabled = enabled
        self.ms = {}
        self._t = None

    def start(self):
        if self.enabled:
            torch.cuda.synchronize()
            self._t = time.perf_counter()

    def lap(self, name):
        if self.enabled:
            torch.cuda.synchronize()
            now = time.perf_counter()
            self.ms[name] = self.ms.get(name, 0.0) + (now - self._t) * 1e3
            self._t = now


_SIDE_STREAM = None


def side_stream():
    """second CUDA stream of the pipeline: the dual-cell passes (needed only by the contouring at the very end) run
    there, next to the grid / search / network stages — all of them latency-bound kernels that leave most of the GPU idle"""
    global _SIDE_STREAM
    if _SIDE_STREAM is None:
        _SIDE_STREAM = torch.cuda.Stream()
    return _SIDE_STREAM


class AsyncDuals:
    """Dual vertex indices computed on the side stream; `get()` joins the streams and returns the tensor."""

    def __init__(self, tree, enabled=True):
        self.tree, self.out = tree, None
        self.enabled = enabled and hasattr(tree, "dual_vertex_indices_begin")
        if self.enabled:
            self.main = torch.cuda.current_stream()
            self.side = side_stream()
            self.side.wait_stream(self.main)  # the octree was built on the main stream
            with torch.cuda.stream(self.side):
                tree.dual_vertex_indices_begin()

    def fill(self):
        """after some main-stream work was queued: wait for the count (done long ago), queue the fill"""
        if self.enabled and self.out is None:
            with torch.cuda.stream(self.side):
                self.out = self.tree.dual_vertex_indices_finish()

    def get(self):
        if not self.enabled:
            if self.out is None:
                self.out = self.tree.dual_vertex_indices()
            return self.out
        self.fill()
        self.main.wait_stream(self.side)
        self.out.record_stream(self.main)
        self.tree.dual_check()
        return self.out


def build_input_dict(points, normals, radii, bb_min, bb_max, levels=5, radius_scale=1.0, max_depth=21, timer=None,
                     K=ops, normals_ready=None, async_duals=False):
    """Grid building + aggregation search (asr.cpp:143-312).  Returns
    (input_dict, dual_vertex_indices, octree).  K = kernel namespace (ops or
    shard.ShardedOps: there the aggregation arrays cover this rank's voxels).
    normals_ready: CUDA event after which `normals` (copied on another stream) may be read."""
    timer = timer or StageTimer()
    timer.start()
    tree = K.Octree(points, radii, bb_min, bb_max, radius_scale, 0, max_depth)
    timer.lap("octree")
    if async_duals and not timer.enabled:
        duals = AsyncDuals(tree)  # runs beside the next stages; the caller joins with duals.get()
    else:
        duals = tree.dual_vertex_indices()
    timer.lap("duals")
    grids = tree.grids(levels, True)
    timer.lap("grids")
    if isinstance(duals, AsyncDuals):
        duals.fill()
    ones = torch.ones((points.shape[0], 1), dtype=torch.float32, device=points.device)
    if normals_ready is not None:
        torch.cuda.current_stream().wait_event(normals_ready)
    d = {"points": points, "feats": torch.cat([normals, ones], 1)}
    for i, g in enumerate(grids):
        for k, v in g.items():
            if k != "voxel_keys":
                d[k + str(i)] = v
    plans_async = (K is ops and ops.SPARSE_CONV_BACKEND == "gx" and not timer.enabled and len(grids) == levels
                   and "neighbors_row_splits0" in d)
    if plans_async:  # the gx plans of the U-Net's tables are built beside the search, on the side stream
        from . import gx
        gx.begin_plans_async(d, levels, side_stream())
    if "voxel_centers0" not in d:  # empty tree
        d["voxel_centers0"] = torch.zeros((0, 3), dtype=torch.float32, device=points.device)
        d["voxel_sizes0"] = torch.zeros(0, dtype=torch.float32, device=points.device)
    idx, dist, rs = K.multi_radius_search(points, d["voxel_centers0"], d["voxel_sizes0"], frame=tree.search_frame())
    d["aggregation_neighbors_index"] = idx
    d["aggregation_neighbors_dist"] = dist
    d["aggregation_row_splits"] = rs
    d["aggregation_scale_compat"] = K.scale_compatibility(d["voxel_sizes0"], radii, idx, rs)
    if plans_async:
        gx.finish_plans_async(d)
    timer.lap("search")
    return d, duals, tree


def run_network(model, input_dict, timer=None):
    """aggregate -> unet -> decode(shift = 0) with channel 0 rescaled by the voxel
    size (asr.cpp:314-336).  Returns values [V0, 2]."""
    timer = timer or StageTimer()
    timer.start()
    feats = model.aggregate(input_dict)
    timer.lap("aggregate")
    code = model.unet(feats, input_dict)
    timer.lap("unet")
    values = model.decode(None, code, signed_scale=input_dict["voxel_sizes0"])
    timer.lap("decode")
    return values


def reconstruct_vertices(model, points, normals, radii, bb_min=None, bb_max=None, levels=None, radius_scale=1.0,
                         max_depth=21, contouring_value_threshold=1.0, timer=None, triangles=False,
                         normals_ready=None):
    """Whole path on device tensors.  bb defaults to the exact min/max of the
    points like the C++ driver (asr.cpp:148-150).  triangles=True also runs the polygon
    passes of CreateTriangleMesh: the result then has "triangles"
    and "vertices" includes the fan-centre vertices after the dual-cell vertices."""
    timer = timer or StageTimer()
    levels = levels or model.octree_levels
    if bb_min is None:
        bb_min = points.min(0).values.cpu().numpy()
        bb_max = points.max(0).values.cpu().numpy()
    K = getattr(model, "K", ops)
    d, duals, tree = build_input_dict(points, normals, radii, bb_min, bb_max, levels, radius_scale, max_depth, timer, K,
                                      normals_ready, async_duals=K is ops)
    values = run_network(model, d, timer)
    if isinstance(duals, AsyncDuals):
        duals = duals.get()
    timer.start()
    tris = None
    if triangles:
        verts, tris, vdual = K.contour_mesh(values, duals, d["voxel_centers0"], contouring_value_threshold)
    else:
        verts, vdual = K.contour_vertices(values, duals, d["voxel_centers0"], contouring_value_threshold)
    timer.lap("contour")
    out = {"vertices": verts, "vertex_dual": vdual, "values": values, "dual_vertex_indices": duals,
           "input_dict": d, "octree": tree}
    if tris is not None:
        out["triangles"] = tris
    return out


_PINNED = {}


def _to_host(t, key):
    """D2H through a cached pinned staging buffer (grown on demand)."""
    n = t.numel()
    buf = _PINNED.get(key)
    if buf is None or buf.numel() < n or buf.dtype != t.dtype:
        buf = torch.empty(max(n, 1) * 5 // 4, dtype=t.dtype, pin_memory=True)
        _PINNED[key] = buf
    view = buf[:n].view(t.shape)
    view.copy_(t, non_blocking=True)
    return view


_COPY_STREAM = None


def _as_host_tensor(a):
    """numpy array or CPU tensor -> contiguous float32 CPU tensor (pinned memory is used as is)."""
    if isinstance(a, torch.Tensor):
        return a.contiguous() if a.dtype == torch.float32 else a.float().contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, np.float32))


def reconstruct_vertices_host(model, points, normals, radii, bb_min=None, bb_max=None, pinned_out=False, shard_ctx=None,
                              **kw):
    """Same path with HOST buffers in and out (numpy arrays or CPU tensors, ideally pinned): H2D of
    the cloud, the device path, D2H of the vertices and SDF values.  This is the call the `e2e`
    benchmark number times.  The normals (first needed by the aggregation, after the octree and the
    grids are built) are copied on a second stream, behind the geometry stages.
    pinned_out=True returns views of the cached pinned staging buffers instead of fresh numpy copies."""
    global _COPY_STREAM
    dev = torch.device("cuda")
    hp, hn, hr = _as_host_tensor(points), _as_host_tensor(normals), _as_host_tensor(radii)
    if bb_min is None:
        bb_min, bb_max = hp.min(0).values.numpy(), hp.max(0).values.numpy()
    main = torch.cuda.current_stream()
    if _COPY_STREAM is None:
        _COPY_STREAM = torch.cuda.Stream()
    p = hp.to(dev, non_blocking=True)
    r = hr.to(dev, non_blocking=True)
    _COPY_STREAM.wait_stream(main)
    with torch.cuda.stream(_COPY_STREAM):
        n = hn.to(dev, non_blocking=True)
        ready = torch.cuda.Event()
        ready.record(_COPY_STREAM)
    n.record_stream(main)
    if shard_ctx is not None:  # multi-GPU: every rank uploads the cloud and runs its share (shard_gx.py)
        from . import shard_gx
        main.wait_event(ready)
        out = shard_gx.reconstruct_vertices(model, shard_ctx, p, n, r, bb_min, bb_max, **kw)
    else:
        out = reconstruct_vertices(model, p, n, r, bb_min, bb_max, normals_ready=ready, **kw)
    v, s = _to_host(out["vertices"], "vertices"), _to_host(out["values"], "values")
    t = _to_host(out["triangles"], "triangles") if "triangles" in out else None
    main.synchronize()
    if pinned_out:
        res = {"vertices": v, "values": s}
        if t is not None:
            res["triangles"] = t
        return res
    res = {"vertices": v.numpy().copy(), "values": s.numpy().copy()}
    if t is not None:
        res["triangles"] = t.numpy().copy()
    return res
