"""Multi-GPU form of the hot path on the gx kernels: one process per GPU, `torch.distributed` (NCCL) for the two
small reductions, a SYMMETRIC-MEMORY arena + peer stores over NVLink / NVSwitch for the halo rows.

Decomposition:

  * The domain is cut along the Z-curve into `world` regions holding the same number of level-0 voxels; on EVERY grid
    level a rank owns the voxels whose Z-curve position (Morton code brought to depth 21) falls into its region, so
    the fine and coarse voxels of one place live on the same rank and only region borders are exchanged.  Ownership
    is a function of the location code alone: every rank evaluates it locally (csrc/shard.cu), nothing is negotiated.
  * The geometry (octree, neighbour tables, dual cells) is built by every rank — it is the input of the ownership
    function and of the "who reads my rows" masks; sharding it is the next step.
  * Search, aggregation, every sparse convolution, and the decoder run on the rank's own rows only: the gx plans are
    built for the owned rows of each table (`gx.Plan(rows=...)`), outputs go to full-size buffers that live at the same
    offset of a symmetric allocation on every rank.
  * After a convolution the rank PUSHES the rows it owns that other ranks' plans read (one ring of face neighbours
    along the region border) straight into those ranks' buffers with peer stores (asr_shard_push), followed by a
    device-side barrier of the symmetric-memory handle; there is no packing, no host synchronisation and no collective
    call on the data path.  Levels below `min_rows` rows are still computed by their owners, but broadcast to everybody.
  * Two NCCL all-reduces per pass remain: the per-voxel pair counts and the first V0 pair importances that the first
    encoder block reads through the reference's quirk 0.

Results are identical to the single-GPU path (same kernels, same per-row arithmetic; the convolutions are
bit-reproducible), checked by bench.py (`parity_vs_1gpu`) and tests/test_gpu_multi.py on real GPUs.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import gx, ops
from ._lib import check, lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Arena:
    """One symmetric allocation per rank (same size everywhere), carved by a bump allocator: the k-th allocation of a
    pass has the same offset on every rank because every rank allocates the same sequence of sizes (the geometry is
    replicated).  `reset()` starts a new pass behind a cross-rank barrier."""

    def __init__(self, nbytes, group):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.nbytes = int(nbytes)
        self.buf = symm_mem.empty(self.nbytes, dtype=torch.uint8, device=torch.device("cuda", torch.cuda.current_device()))
        self.hdl = symm_mem.rendezvous(self.buf, group)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.peers = (C.c_void_p * self.world)(*ptrs)
        self.base = self.buf.data_ptr()
        self.off = 0
        self.pushed_bytes = 0

    def reset(self):
        self.off = 0
        self.barrier()

    def barrier(self):
        self.hdl.barrier(channel=0)

    def alloc(self, shape, dtype):
        n = int(np.prod(shape)) * torch.empty(0, dtype=dtype).element_size()
        start = (self.off + 255) // 256 * 256
        if start + n > self.nbytes:
            raise RuntimeError("asr_b200 shard arena exhausted: need %d more bytes (ASR_SHARD_ARENA_GB)" % (start + n - self.nbytes))
        self.off = start + n
        self.peak = max(getattr(self, "peak", 0), self.off)
        return self.buf[start:start + n].view(dtype).view(*shape)

    def offset_of(self, t):
        return t.data_ptr() - self.base


class ShardContext(gx.LocalContext):
    """gx buffer / halo context of one rank (see the module docstring)."""

    def __init__(self, arena, min_rows=16384):
        self.arena = arena
        self.rank, self.world = arena.rank, arena.world
        self.min_rows = min_rows
        self.levels = None
        self.exchanges = 0

    # ------------------------------------------------------------------ ownership
    def prepare(self, keys, input_dict, levels):
        """keys[l]: int64 location codes of grid level l.  Computes region thresholds, owners, owned row lists and
        the per-level masks of ranks that read each owned row."""
        L = lib()
        dev = keys[0].device
        V = [int(k.shape[0]) for k in keys]
        pos0 = torch.empty(V[0], dtype=torch.int64, device=dev)
        check(L.asr_shard_positions(_ptr(keys[0]), V[0], _ptr(pos0), _stream()))
        # positions are < 2^63, so the signed sort orders them like the unsigned values
        srt = torch.sort(pos0).values
        cut = torch.tensor([(k * V[0]) // self.world for k in range(1, self.world)], dtype=torch.int64, device=dev)
        thr = srt[cut].contiguous() if self.world > 1 else torch.empty(0, dtype=torch.int64, device=dev)
        self.owner, self.rows, self.need = [], [], []
        for l in range(levels):
            o = torch.empty(V[l], dtype=torch.uint8, device=dev)
            check(L.asr_shard_owner(_ptr(keys[l]), V[l], _ptr(thr), self.world - 1, _ptr(o), _stream()))
            self.owner.append(o)
            self.need.append(torch.zeros(V[l], dtype=torch.int32, device=dev))
        d = input_dict
        for l in range(levels):
            # within-grid table: rows and inputs on level l
            self._need(d["neighbors_row_splits%d" % l], d["neighbors_index%d" % l], l, l)
        self.inv = []
        for l in range(levels - 1):
            ui, uk, us = d["up_neighbors_index%d" % l], d["up_neighbors_kernel_index%d" % l], d["up_neighbors_row_splits%d" % l]
            self._need(us, ui, l, l + 1)  # up table: rows on level l read level l + 1
            inv = ops.invert_neighbors_list(V[l + 1], ui, us, uk)
            self.inv.append(inv)
            self._need(inv.neighbors_row_splits, inv.neighbors_index, l + 1, l)  # down table: rows on l + 1 read l
        self.border = []
        for l in range(levels):
            rows = torch.nonzero(self.owner[l] == self.rank).reshape(-1).to(torch.int32)
            self.rows.append(rows)
            if V[l] < self.min_rows:
                self.need[l] = None  # small level: every owned row goes to every rank
                self.border.append(rows)
            else:
                # the owned rows somebody else reads (the region border): the only rows a push has to look at
                self.border.append(rows[self.need[l][rows.long()] != 0].contiguous())
        self.V = V
        self.levels = levels

    def _need(self, row_splits, idx, out_level, in_level):
        check(lib().asr_shard_need_mask(_ptr(row_splits), _ptr(idx), row_splits.shape[0] - 1, _ptr(self.owner[out_level]),
                                        _ptr(self.owner[in_level]), self.rank, _ptr(self.need[in_level]), _stream()))

    # ------------------------------------------------------------------ gx context
    def empty(self, V, C_, device):
        buf = self.arena.alloc((V + 1, 2 * C_), torch.float16)
        buf[V:].zero_()
        return gx.H2(buf, C_, 0, C_)

    def mark(self):
        return self.arena.off

    def release(self, mark):
        # buffers allocated after `mark` were temporaries of a block: every rank has passed the barrier that followed
        # the last push into them before anything is allocated there again (the next convolution's output)
        self.arena.off = mark
        self.arena.peak = max(getattr(self.arena, "peak", 0), mark)

    def plans(self, input_dict, levels):
        cache = input_dict.get("_asr_gx_shard_plans")
        if cache is not None:
            return cache
        d, V = input_dict, self.V
        P = {"nb": [], "up": [], "down": [], "V": V}
        for i in range(levels):
            P["nb"].append(gx.Plan(d["neighbors_index%d" % i], d["neighbors_kernel_index%d" % i],
                                   d["neighbors_row_splits%d" % i], V[i], 55, gx.MODE_STATIONARY, rows=self.rows[i]))
        for i in range(levels - 1):
            ui, uk, us = d["up_neighbors_index%d" % i], d["up_neighbors_kernel_index%d" % i], d["up_neighbors_row_splits%d" % i]
            P["up"].append(gx.Plan(ui, uk, us, V[i + 1], 9, gx.MODE_PAIR_FINAL, rows=self.rows[i]))
            inv = self.inv[i]
            P["down"].append(gx.Plan(inv.neighbors_index, inv.neighbors_attributes, inv.neighbors_row_splits, V[i], 9,
                                     gx.MODE_STATIONARY, rows=self.rows[i + 1]))
        for group in ("nb", "up", "down"):
            for p in P[group]:
                p.finish()
        input_dict["_asr_gx_shard_plans"] = P
        return P

    def push_rows(self, tensor, level, seg_off, seg_len, all_rows=None):
        """copies this rank's rows of `tensor` (a view into the arena, row-major 2-D) that other ranks read into
        their copies; then a cross-rank barrier on the stream."""
        if self.world > 1:
            rows = self.border[level] if all_rows is None else all_rows
            mask = self.need[level] if all_rows is None else None
            segs = (C.c_int * 2)(*(list(seg_off) + [0])[:2])
            pitch = tensor.stride(0) * tensor.element_size()
            check(lib().asr_shard_push(self.arena.peers, self.world, self.rank, self.arena.offset_of(tensor), pitch, segs,
                                       len(seg_off), int(seg_len), _ptr(rows), rows.shape[0], _ptr(mask), _stream()))
            self.arena.barrier()
            self.exchanges += 1

    def done(self, view, level):
        # hi and lo planes of the view's channels
        self.push_rows(view.buf, level, (view.hi * 2, view.lo * 2), view.C * 2)


def first_pair_importances(rs, own_rows, imp_pairs, V0, all_reduce=None):
    """Quirk 0 on a sharded level 0: the first encoder block reads imp_pairs[v], v < V0, of the GLOBAL
    pair list (all voxels' pairs in table order).  rs / imp_pairs: row splits and pair importances of this rank's own
    voxels `own_rows` (ascending table rows); all_reduce(tensor): in-place sum over the ranks (None: one rank).
    Returns (first [V0] float32, total number of pairs).  Only the few voxels whose pairs fall below V0 are touched
    (a device-only formulation over ALL local pairs was measured slower: 5.0 vs 3.7 ms at 2 GPUs, 10 M points)."""
    dev = imp_pairs.device
    counts = torch.zeros(V0, dtype=torch.int32, device=dev)
    counts[own_rows] = (rs[1:] - rs[:-1]).to(torch.int32)
    if all_reduce is not None:
        all_reduce(counts)
    start = torch.cumsum(counts.long(), 0) - counts.long()  # first global pair of every voxel
    total = int(start[-1] + counts[-1]) if V0 else 0
    first = torch.zeros(V0, dtype=torch.float32, device=dev)
    s_own = start[own_rows]
    sel = torch.nonzero(s_own < V0).reshape(-1)  # owned voxels whose pairs can fall into [0, V0)
    if sel.numel():
        lo = rs[sel]
        n = torch.minimum(rs[sel + 1] - lo, V0 - s_own[sel])
        tot = int(n.sum())
        if tot:
            seg = torch.repeat_interleave(torch.arange(sel.numel(), device=dev), n)
            within = torch.arange(tot, device=dev) - torch.repeat_interleave(torch.cumsum(n, 0) - n, n)
            first[s_own[sel][seg] + within] = imp_pairs[lo[seg] + within]
    if all_reduce is not None:
        all_reduce(first)
    return first, total


def reconstruct_vertices(net, ctx, points, normals, radii, bb_min, bb_max, levels=None, radius_scale=1.0, max_depth=21,
                         contouring_value_threshold=1.0, timer=None):
    """pipeline.reconstruct_vertices on `ctx.world` GPUs (called by every rank with the same cloud).  Returns the same
    dict; `values`, `vertices`, `vertex_dual` are complete on every rank."""
    from . import pipeline
    timer = timer or pipeline.StageTimer()
    levels = levels or net.octree_levels
    rank, world = ctx.rank, ctx.world
    ctx.arena.reset()
    timer.start()
    tree = ops.Octree(points, radii, bb_min, bb_max, radius_scale, 0, max_depth)
    timer.lap("octree")
    duals = pipeline.AsyncDuals(tree, enabled=not timer.enabled)  # side stream, joined before the contouring
    if timer.enabled:
        duals.get()
    timer.lap("duals")
    grids = tree.grids(levels, True)
    timer.lap("grids")
    duals.fill()
    d = {"points": points}
    for i, g in enumerate(grids):
        for k, v in g.items():
            if k != "voxel_keys":
                d[k + str(i)] = v
    keys = [g["voxel_keys"] for g in grids]
    ctx.prepare(keys, d, levels)
    own0 = ctx.rows[0]
    own0_l = own0.long()
    V0 = ctx.V[0]
    timer.lap("shard_prepare")
    # ---- aggregation search + continuous convolution on the owned level-0 voxels
    centers = d["voxel_centers0"].index_select(0, own0_l)
    sizes = d["voxel_sizes0"].index_select(0, own0_l)
    idx, dist2, rs = ops.multi_radius_search(points, centers, sizes, frame=tree.search_frame())
    compat = ops.scale_compatibility(sizes, radii, idx, rs)
    timer.lap("search")
    imp_pairs = ops.aggregation_importance(compat, dist2)
    c = net.cconv_block_in.conv1
    ones = torch.ones((points.shape[0], 1), dtype=torch.float32, device=points.device)
    feats_in = torch.cat([normals, ones], 1)
    feats_own = ops.continuous_conv(c.kernel, centers, sizes, c.offset, points, feats_in, None, idx, imp_pairs, rs,
                                    normalize=True, bias=c.bias, relu=True)
    # ---- quirk 0: the first encoder block reads imp_pairs[v] for the GLOBAL pair list (voxels in table order)
    first, pairs_total = first_pair_importances(
        rs, own0_l, imp_pairs, V0, (lambda t: dist.all_reduce(t, group=ctx.arena.group)) if world > 1 else None)
    if pairs_total < V0:
        raise IndexError("fewer aggregation pairs than voxels")
    # ---- features into the symmetric split-half buffer: ONLY the owned rows are written (the other ranks push their
    # rows into this buffer at their own pace), then the halo is pushed
    x0 = ctx.empty(V0, feats_own.shape[1], points.device)
    gx.from_f32(feats_own, out=x0, rows=own0)
    ctx.done(x0, 0)
    d["aggregation_neighbors_index"], d["aggregation_neighbors_dist"], d["aggregation_row_splits"] = idx, dist2, rs
    d["aggregation_scale_compat"] = compat
    d["aggregation_pairs_total"] = pairs_total
    timer.lap("aggregate")
    code = ctx.arena.alloc((V0, 32), torch.float32)
    gx.unet(net, (None, first), d, ctx=ctx, x0=x0, code=code)
    timer.lap("unet")
    # ---- decoder on the owned rows, values to everybody
    values = ctx.arena.alloc((V0, 2), torch.float32)
    v_own = net.decode(None, code.index_select(0, own0_l), signed_scale=sizes)
    values[own0_l] = v_own
    if world > 1:
        _broadcast_values(ctx, values)  # every rank contours the whole field
    timer.lap("decode")
    duals = duals.get()
    verts, vdual = ops.contour_vertices(values, duals, d["voxel_centers0"], contouring_value_threshold)
    timer.lap("contour")
    return {"vertices": verts, "vertex_dual": vdual, "values": values, "dual_vertex_indices": duals, "input_dict": d,
            "octree": tree}


def _broadcast_values(ctx, values):
    """[V0, 2] float32 values: this rank's rows to every rank.  The push kernel moves 16-byte pieces, so the rows go
    through a staging buffer padded to 4 floats per row (one piece per row)."""
    V0 = values.shape[0]
    wide = ctx.arena.alloc((V0, 4), torch.float32)
    own = ctx.rows[0].long()
    wide[own, :2] = values[own]
    ctx.push_rows(wide, 0, (0,), 16, all_rows=ctx.rows[0])
    values.copy_(wide[:, :2])
