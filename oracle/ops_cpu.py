"""TEST INFRASTRUCTURE ONLY — CPU restatement of the float half of the hot path.

PARITY UNPINNED at this boundary: the arithmetic of these ops lives in Open3D
v0.14.1 (pinned by the reference at cmake/external_deps.cmake:84-85) and
nanoflann v1.3.2 (:58), neither of which is vendored in /root/reference, is
installed in this image, or is covered by any reference test.  The functions
below restate the published algorithms as recalled in SURVEY.md §8c and are
anchored on the reference's own call sites:

  multi_radius_search        nsearch.cpp:130-146, datareader.py:776-785
  scale_compatibility        nsearch.cpp:149-161, models/common.py:19-44
  window_poly6               models/common_torch.py:21-22
  continuous_conv            net_definitions_torch.py:59-70,108-116
  sparse_conv                models/common_torch.py:133-142
  reduce_subarrays_sum       models/common_torch.py:127
  invert_neighbors_list      net_definitions_torch.py:30-35

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import
this module; it is the checker, never the product.
"""
from collections import namedtuple

import numpy as np
import torch

# --------------------------------------------------------------------------- search


def multi_radius_search(points, queries, radii):
    """Open3D NearestNeighborSearch.MultiRadiusSearch semantics (nanoflann
    radiusSearch, sorted=True): neighbours with d2 < r*r (strict), ascending by
    d2, where d2 = ((dx*dx) + dy*dy) + dz*dz in float32 without FMA
    (nanoflann L2_Adaptor tail loop for dim 3).  Ties in d2 are ordered by point
    index here (the reference's order for exact ties is unspecified).

    Returns (index int32[P], d2 float32[P], row_splits int64[Q+1]).
    """
    from scipy.spatial import cKDTree

    points = np.ascontiguousarray(points, np.float32)
    queries = np.ascontiguousarray(queries, np.float32)
    radii = np.ascontiguousarray(radii, np.float32)
    tree = cKDTree(points.astype(np.float64))
    cand = tree.query_ball_point(queries.astype(np.float64), radii.astype(np.float64) * 1.001 + 1e-12,
                                 workers=-1, return_sorted=True)
    lens = np.fromiter((len(c) for c in cand), np.int64, len(cand))
    flat = np.fromiter((i for c in cand for i in c), np.int64, int(lens.sum()))
    qid = np.repeat(np.arange(len(cand)), lens)
    d = queries[qid] - points[flat]  # float32
    d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    keep = d2 < (radii * radii)[qid]
    flat, qid, d2 = flat[keep], qid[keep], d2[keep]
    order = np.lexsort((flat, d2, qid))
    flat, qid, d2 = flat[order], qid[order], d2[order]
    row_splits = np.zeros(len(queries) + 1, np.int64)
    np.cumsum(np.bincount(qid, minlength=len(queries)), out=row_splits[1:])
    return flat.astype(np.int32), d2.astype(np.float32), row_splits


def scale_compatibility(voxel_sizes, point_radii, neighbors_index, neighbors_row_splits, gamma=2.0):
    """(min(s_v, 2 r_p) / max(s_v, 2 r_p)) ** gamma per pair (nsearch.cpp:149-161)."""
    voxel_sizes = np.asarray(voxel_sizes, np.float32)
    lens = np.diff(neighbors_row_splits)
    a = np.repeat(voxel_sizes, lens)
    b = np.float32(2) * np.asarray(point_radii, np.float32)[neighbors_index]
    q = np.minimum(a, b) / np.maximum(a, b)
    return np.power(q, np.float32(gamma)).astype(np.float32)


def window_poly6(r_sqr):
    return torch.clamp((1 - r_sqr)**3, 0, 1)


# --------------------------------------------------------------------------- helpers


def _rows(row_splits):
    lens = row_splits[1:] - row_splits[:-1]
    return torch.repeat_interleave(torch.arange(lens.shape[0]), lens), lens


def reduce_subarrays_sum(values, row_splits):
    """out[i] = sum(values[row_splits[i]:row_splits[i+1]]); empty row -> 0."""
    rows, lens = _rows(row_splits)
    out = torch.zeros(lens.shape[0], dtype=values.dtype)
    out.index_add_(0, rows, values)
    return out


InvertResult = namedtuple("InvertResult", ["neighbors_index", "neighbors_row_splits", "neighbors_attributes"])


def invert_neighbors_list(num_points, inp_neighbors_index, inp_neighbors_row_splits, inp_neighbors_attributes):
    """CSR transpose: stable counting sort of the entries by target index."""
    idx = inp_neighbors_index.to(torch.int64)
    rows, _ = _rows(inp_neighbors_row_splits)
    order = torch.argsort(idx, stable=True)
    counts = torch.bincount(idx, minlength=num_points)
    rs = torch.zeros(num_points + 1, dtype=torch.int64)
    rs[1:] = torch.cumsum(counts, 0)
    attrs = inp_neighbors_attributes[order] if inp_neighbors_attributes.numel() else inp_neighbors_attributes
    return InvertResult(rows[order].to(inp_neighbors_index.dtype), rs, attrs)


# --------------------------------------------------------------------------- convolutions


def sparse_conv(filters, inp_features, inp_importance, neighbors_index, neighbors_kernel_index,
                neighbors_importance, neighbors_row_splits, normalize, dtype=None):
    """out[o] = sum_n filters[k_n]^T (imp_n * x[idx_n]); optionally / sum_n imp_n
    (or the neighbour count when no importance is given) where that is != 0."""
    dt = dtype or inp_features.dtype
    W = filters.to(dt)
    idx = neighbors_index.to(torch.int64)
    kidx = neighbors_kernel_index.to(torch.int64)
    rows, lens = _rows(neighbors_row_splits)
    x = inp_features.to(dt)[idx]
    imp = None
    if neighbors_importance.numel():
        imp = neighbors_importance.to(dt)
    if inp_importance.numel():
        pimp = inp_importance.to(dt)[idx]
        imp = pimp if imp is None else imp * pimp
    if imp is not None:
        x = x * imp[:, None]
    out = torch.zeros(lens.shape[0], W.shape[2], dtype=dt)
    for k in range(W.shape[0]):
        sel = torch.nonzero(kidx == k).squeeze(1)
        if sel.numel():
            out.index_add_(0, rows[sel], x[sel] @ W[k])
    if normalize:
        if neighbors_importance.numel():
            norm = torch.zeros(lens.shape[0], dtype=dt).index_add_(0, rows, neighbors_importance.to(dt))
        else:
            norm = lens.to(dt)
        nz = norm != 0
        out[nz] = out[nz] / norm[nz][:, None]
    return out


def filter_coordinates(rel, extents_per_pair, offset, size):
    """ball_to_cube_radial + align_corners mapping of relative positions to
    continuous kernel coordinates in [0, size-1] (before clamping)."""
    x = rel * (2.0 / extents_per_pair)[:, None]
    norm = torch.sqrt((x * x).sum(1))
    amax = x.abs().max(1).values
    s = torch.where(amax < 1e-8, torch.zeros_like(norm), 0.5 * norm / torch.where(amax < 1e-8, torch.ones_like(amax), amax))
    x = x * s[:, None]
    return (x + offset[None, :] + 0.5) * (size - 1)


def continuous_conv(filters, out_positions, extents, offset, inp_positions, inp_features, inp_importance,
                    neighbors_index, neighbors_importance, neighbors_row_splits, normalize=True, dtype=None):
    """Open3D continuous_conv with align_corners=True, coordinate_mapping=
    'ball_to_cube_radial', interpolation='linear' (the configuration the
    reference uses, net_definitions_torch.py:53-57): trilinear splat of
    imp*feat into a [Sz,Sy,Sx,Cin] cell tensor, contraction with the filter,
    division by sum(imp) where != 0.  filters layout [Sz,Sy,Sx,Cin,Cout]."""
    dt = dtype or inp_features.dtype
    Sz, Sy, Sx, Cin, Cout = filters.shape
    idx = neighbors_index.to(torch.int64)
    rows, lens = _rows(neighbors_row_splits)
    V = lens.shape[0]
    ext = extents.to(dt)
    if ext.numel() == 1:
        ext = ext.reshape(1).expand(V)
    rel = inp_positions.to(dt)[idx] - out_positions.to(dt)[rows]
    size = torch.tensor([Sx, Sy, Sz], dtype=dt)
    c = filter_coordinates(rel, ext[rows], offset.to(dt), size)
    c = torch.minimum(torch.clamp(c, min=0), (size - 1)[None, :])
    i0 = torch.minimum(c.floor(), (size - 1)[None, :]).to(torch.int64)
    i1 = torch.minimum(i0 + 1, (size - 1).to(torch.int64)[None, :])
    a = torch.clamp(c - i0.to(dt), 0, 1)
    imp = torch.ones(idx.shape[0], dtype=dt)
    if neighbors_importance.numel():
        imp = neighbors_importance.to(dt)
    norm = torch.zeros(V, dtype=dt).index_add_(0, rows, imp)
    if inp_importance.numel():
        imp = imp * inp_importance.to(dt)[idx]
    f = inp_features.to(dt)[idx] * imp[:, None]
    B = torch.zeros(V, Sz * Sy * Sx, Cin, dtype=dt)
    Bf = B.view(-1, Cin)
    for tz in (0, 1):
        wz = a[:, 2] if tz else 1 - a[:, 2]
        iz = i1[:, 2] if tz else i0[:, 2]
        for ty in (0, 1):
            wy = a[:, 1] if ty else 1 - a[:, 1]
            iy = i1[:, 1] if ty else i0[:, 1]
            for tx in (0, 1):
                wx = a[:, 0] if tx else 1 - a[:, 0]
                ix = i1[:, 0] if tx else i0[:, 0]
                w = wx * wy * wz
                cell = (iz * Sy + iy) * Sx + ix
                Bf.index_add_(0, rows * (Sz * Sy * Sx) + cell, f * w[:, None])
    out = B.view(V, -1) @ filters.to(dt).reshape(-1, Cout)
    if normalize:
        nz = norm != 0
        out[nz] = out[nz] / norm[nz][:, None]
    return out


# ---------------------------------------------------------------- point pre-processing (row f-2)
def _f32_d2(q, p):
    """nanoflann L2_Simple_Adaptor for dim 3 in float32: ((dx*dx) + dy*dy) + dz*dz."""
    d = (q.astype(np.float32) - p.astype(np.float32)).astype(np.float32)
    return ((d[..., 0] * d[..., 0]).astype(np.float32) + (d[..., 1] * d[..., 1]).astype(np.float32)).astype(
        np.float32) + (d[..., 2] * d[..., 2]).astype(np.float32)


def knn(points, k, extra=8):
    """(sorted float32 squared distances [N, k], indices [N, k]) of the k nearest points (the point
    itself included), reference KDTree::ComputeKRadius / ComputeInlier (nsearch.cpp:30-85).  The
    candidates come from a float64 cKDTree query of k + extra neighbours, the ranking is redone
    with nanoflann's float32 arithmetic."""
    from scipy.spatial import cKDTree
    points = np.ascontiguousarray(points, np.float32)
    kk = min(points.shape[0], k + extra)
    _, idx = cKDTree(points.astype(np.float64)).query(points.astype(np.float64), k=kk, workers=-1)
    idx = idx.reshape(points.shape[0], kk)
    d2 = _f32_d2(points[:, None, :], points[idx])
    order = np.argsort(d2, axis=1, kind="stable")[:, :min(k, kk)]
    return np.take_along_axis(d2, order, 1), np.take_along_axis(idx, order, 1)


def k_radius(points, k):
    d2, _ = knn(points, k)
    return np.sqrt(d2.max(1)).astype(np.float32)


def knn_inlier(points, radii, radius_fraction=0.5, k=24, outlier_threshold=1):
    _, idx = knn(points, k)
    radii = np.asarray(radii, np.float32)
    votes = (radii[idx] < (radii * np.float32(radius_fraction)).astype(np.float32)[:, None]).sum(1)
    return votes < outlier_threshold


def radius_neighbor_counts(points, radii):
    """number of points with d2 < r^2 in float32 (ComputeRadiusNeighbors, nsearch.cpp:87-105)."""
    from scipy.spatial import cKDTree
    points = np.ascontiguousarray(points, np.float32)
    radii = np.asarray(radii, np.float32)
    tree = cKDTree(points.astype(np.float64))
    out = np.zeros(points.shape[0], np.int32)
    cand = tree.query_ball_point(points.astype(np.float64), radii.astype(np.float64) * 1.001 + 1e-12, workers=-1)
    for i, c in enumerate(cand):
        c = np.asarray(c, np.int64)
        out[i] = int((_f32_d2(points[i][None, :], points[c]) < np.float32(radii[i]) * np.float32(radii[i])).sum())
    return out
