"""Oracle-backed stand-ins for the four Open3D-ML torch ops the reference calls
(SURVEY.md §8b B2), with Open3D v0.14.1's full op schemas (defaulted trailing arguments included).
Registered under the private namespace `open3d_oracle::` so they can coexist with the product's
`open3d::` CUDA ops in one process; `ASR_ORACLE_O3D_NAMESPACE=open3d` (set by
tests/golden/make_traced_archive.py, a process that never imports the product shim) registers them
as `open3d::*` instead, so that an archive traced here carries the op names real Open3D would."""
import os
from typing import NamedTuple

import torch

from oracle import ops_cpu as _o

NAMESPACE = os.environ.get("ASR_ORACLE_O3D_NAMESPACE", "open3d_oracle")
_lib = torch.library.Library(NAMESPACE, "DEF")
_lib.define("invert_neighbors_list(int num_points, Tensor inp_neighbors_index, Tensor inp_neighbors_row_splits, "
            "Tensor inp_neighbors_attributes) -> (Tensor neighbors_index, Tensor neighbors_row_splits, "
            "Tensor neighbors_attributes)")
_lib.define("reduce_subarrays_sum(Tensor values, Tensor row_splits) -> Tensor")
_lib.define("sparse_conv(Tensor filters, Tensor inp_features, Tensor inp_importance, Tensor neighbors_index, "
            "Tensor neighbors_kernel_index, Tensor neighbors_importance, Tensor neighbors_row_splits, "
            "bool normalize=False, int max_temp_mem_MB=64) -> Tensor")
_lib.define("continuous_conv(Tensor filters, Tensor out_positions, Tensor extents, Tensor offset, "
            "Tensor inp_positions, Tensor inp_features, Tensor inp_importance, Tensor neighbors_index, "
            "Tensor neighbors_importance, Tensor neighbors_row_splits, bool align_corners=False, "
            "str coordinate_mapping=\"ball_to_cube_radial\", bool normalize=False, "
            "str interpolation=\"linear\", int max_temp_mem_MB=64) -> Tensor")
_ops = getattr(torch.ops, NAMESPACE)


def _inv(num_points, idx, rs, attrs):
    r = _o.invert_neighbors_list(num_points, idx, rs, attrs)
    return r.neighbors_index, r.neighbors_row_splits, r.neighbors_attributes


def _sparse_conv(filters, x, inp_importance, idx, kidx, nimp, rs, normalize=False, max_temp_mem_MB=64):
    return _o.sparse_conv(filters, x, inp_importance, idx, kidx, nimp, rs, normalize)


def _cconv(filters, out_positions, extents, offset, inp_positions, inp_features, inp_importance, idx, nimp, rs,
           align_corners=False, coordinate_mapping="ball_to_cube_radial", normalize=False, interpolation="linear",
           max_temp_mem_MB=64):
    if not (align_corners and coordinate_mapping == "ball_to_cube_radial" and interpolation == "linear"):
        raise NotImplementedError("oracle restates only the configuration the reference uses")
    return _o.continuous_conv(filters, out_positions, extents, offset, inp_positions, inp_features, inp_importance,
                              idx, nimp, rs, normalize=normalize)


_lib.impl("invert_neighbors_list", _inv, "CPU")
_lib.impl("reduce_subarrays_sum", _o.reduce_subarrays_sum, "CPU")
_lib.impl("sparse_conv", _sparse_conv, "CPU")
_lib.impl("continuous_conv", _cconv, "CPU")


class InvertNeighborsListResult(NamedTuple):
    neighbors_index: torch.Tensor
    neighbors_row_splits: torch.Tensor
    neighbors_attributes: torch.Tensor


if NAMESPACE == "open3d":

    def invert_neighbors_list(num_points: int, inp_neighbors_index: torch.Tensor, inp_neighbors_row_splits: torch.Tensor,
                              inp_neighbors_attributes: torch.Tensor):
        a, b, c = torch.ops.open3d.invert_neighbors_list(num_points, inp_neighbors_index, inp_neighbors_row_splits,
                                                         inp_neighbors_attributes)
        return InvertNeighborsListResult(a, b, c)
else:

    def invert_neighbors_list(num_points: int, inp_neighbors_index: torch.Tensor, inp_neighbors_row_splits: torch.Tensor,
                              inp_neighbors_attributes: torch.Tensor):
        a, b, c = torch.ops.open3d_oracle.invert_neighbors_list(num_points, inp_neighbors_index,
                                                                inp_neighbors_row_splits, inp_neighbors_attributes)
        return InvertNeighborsListResult(a, b, c)


def reduce_subarrays_sum(values, row_splits):
    return _ops.reduce_subarrays_sum(values, row_splits)


def sparse_conv(filters, inp_features, inp_importance, neighbors_index, neighbors_kernel_index,
                neighbors_importance, neighbors_row_splits, normalize=False, max_temp_mem_MB=64):
    return _ops.sparse_conv(filters, inp_features, inp_importance, neighbors_index, neighbors_kernel_index,
                            neighbors_importance, neighbors_row_splits, normalize, max_temp_mem_MB)


def continuous_conv(filters, out_positions, extents, offset, inp_positions, inp_features, inp_importance,
                    neighbors_index, neighbors_importance, neighbors_row_splits, align_corners=False,
                    coordinate_mapping="ball_to_cube_radial", normalize=False, interpolation="linear",
                    max_temp_mem_MB=64):
    return _ops.continuous_conv(filters, out_positions, extents, offset, inp_positions, inp_features,
                                inp_importance, neighbors_index, neighbors_importance, neighbors_row_splits,
                                align_corners, coordinate_mapping, normalize, interpolation, max_temp_mem_MB)
