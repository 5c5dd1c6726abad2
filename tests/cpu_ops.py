"""TEST ONLY — a CPU kernel namespace with the same call surface as asr_b200.ops,
backed by the oracle, so the host-side logic that sits above the kernels (model
sequencing, pipeline, multi-GPU sharding in shard.ShardedOps) can be exercised
without a GPU (gloo, world_size 2)."""
from collections import namedtuple

import numpy as np
import torch

from oracle import geomlib, model_cpu, ops_cpu

SPARSE_CONV_BACKEND = "fp32"
PackedFilters = None


class Octree:
    def __init__(self, points, radii, bb_min, bb_max, radius_scale=1.0, grow_steps=0, max_depth=21):
        self._t = geomlib.PortOctree(points.numpy(), radii.numpy(), bb_min, bb_max, radius_scale, grow_steps, max_depth)

    def grids(self, num_levels, voxel_info_all_levels=False):
        out = []
        for g in self._t.grids(num_levels, voxel_info_all_levels):
            out.append({k: torch.from_numpy(v.view(np.int64) if k == "voxel_keys" else v) for k, v in g.items()})
        return out

    def dual_vertex_indices(self):
        return torch.from_numpy(self._t.dual_vertex_indices().astype(np.int64))

    def search_frame(self):
        return None


class ConvPlan:
    def __init__(self, idx, slot, rs, kernel_size):
        self.idx, self.slot, self.row_splits, self.kernel_size = idx, slot, rs, kernel_size
        self.num_out = rs.shape[0] - 1


def sparse_conv(plan, filters, x, inp_importance=None, neighbors_importance=None, importance_col=0, normalize=False,
                normalize_col=0, normalizer=None, bias=None, relu=False, out=None, backend=None):
    e = torch.empty(0)
    idx, slot, rs = plan.idx, plan.slot, plan.row_splits
    plain = ops_cpu.sparse_conv(filters, x, e, idx, slot, e, rs, False)
    res = plain
    imp = None
    if inp_importance is not None:
        imp = inp_importance[idx.long()]
    if neighbors_importance is not None:
        imp = neighbors_importance if imp is None else imp * neighbors_importance
    if imp is not None:
        weighted = ops_cpu.sparse_conv(filters, x, e, idx, slot, imp, rs, False)
        res = torch.cat([plain[:, :importance_col], weighted[:, importance_col:]], 1)
    if normalize:
        nrm = normalizer if normalizer is not None else (rs[1:] - rs[:-1]).float()
        nz = nrm != 0
        res = res.clone()
        res[nz, normalize_col:] = res[nz, normalize_col:] / nrm[nz][:, None]
    if bias is not None:
        res = res + bias
    if relu:
        res = torch.relu(res)
    if out is not None:
        out.copy_(res)
        return out
    return res


def cat(tensors, dim=-1):
    return torch.cat(tensors, dim)


def add(x, y):
    return x + y


def reduce_subarrays_sum(values, row_splits, index=None):
    v = values if index is None else values[index.long()]
    return ops_cpu.reduce_subarrays_sum(v, row_splits)


def invert_neighbors_list(num_points, idx, rs, attrs):
    return ops_cpu.invert_neighbors_list(num_points, idx, rs, attrs)


def multi_radius_search(points, queries, radii, frame=None):
    i, d, r = ops_cpu.multi_radius_search(points.numpy(), queries.numpy(), radii.numpy())
    return torch.from_numpy(i), torch.from_numpy(d), torch.from_numpy(r)


def scale_compatibility(voxel_sizes, point_radii, idx, rs):
    return torch.from_numpy(ops_cpu.scale_compatibility(voxel_sizes.numpy(), point_radii.numpy(), idx.numpy(), rs.numpy()))


def aggregation_importance(compat, dist):
    return compat * ops_cpu.window_poly6(dist)


def pair_importance_for_unet(importance, num_voxels):
    return importance


def continuous_conv(filters, out_positions, extents, offset, inp_positions, inp_features, inp_importance, idx, nimp, rs,
                    normalize=True, bias=None, relu=False, out=None):
    e = torch.empty(0)
    res = ops_cpu.continuous_conv(filters, out_positions, extents, offset if offset is not None else torch.zeros(3),
                                  inp_positions, inp_features, e if inp_importance is None else inp_importance, idx,
                                  e if nimp is None else nimp, rs, normalize=normalize)
    if bias is not None:
        res = res + bias
    if relu:
        res = torch.relu(res)
    if out is not None:
        out.copy_(res)
        return out
    return res


def decode(shifts, code, w1, b1, w2, b2, w3, signed_scale=None, with_gradient=False):
    P = {"dense_decoder1.weight": w1, "dense_decoder1.bias": b1, "dense_decoder2.weight": w2,
         "dense_decoder2.bias": b2, "dense_decoder3.weight": w3}
    s = torch.zeros(code.shape[0], 3) if shifts is None else shifts
    if with_gradient:
        return model_cpu.decode_with_gradient(P, s, code)
    v = model_cpu.decode(P, s, code).clone()
    if signed_scale is not None:
        v[:, 0] *= signed_scale
    return v


def contour_vertices(values, duals, positions, threshold=1.0):
    v, d = geomlib.contour_vertices(values.numpy(), duals.numpy().astype(np.uint64), positions.numpy(), threshold)
    return torch.from_numpy(v), torch.from_numpy(d.astype(np.int64))
