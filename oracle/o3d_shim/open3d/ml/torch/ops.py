"""Oracle-backed stand-ins for the four Open3D-ML torch ops the reference calls
(SURVEY.md §8b B2).  Registered under the private namespace `open3d_oracle::`
so they can coexist with the product's `open3d::` CUDA ops in one process."""
from typing import NamedTuple

import torch

from oracle import ops_cpu as _o

_lib = torch.library.Library("open3d_oracle", "DEF")
_lib.define("invert_neighbors_list(int num_points, Tensor inp_neighbors_index, Tensor inp_neighbors_row_splits, "
            "Tensor inp_neighbors_attributes) -> (Tensor, Tensor, Tensor)")
_lib.define("reduce_subarrays_sum(Tensor values, Tensor row_splits) -> Tensor")
_lib.define("sparse_conv(Tensor filters, Tensor inp_features, Tensor inp_importance, Tensor neighbors_index, "
            "Tensor neighbors_kernel_index, Tensor neighbors_importance, Tensor neighbors_row_splits, "
            "bool normalize) -> Tensor")
_lib.define("continuous_conv(Tensor filters, Tensor out_positions, Tensor extents, Tensor offset, "
            "Tensor inp_positions, Tensor inp_features, Tensor inp_importance, Tensor neighbors_index, "
            "Tensor neighbors_importance, Tensor neighbors_row_splits, bool normalize) -> Tensor")


def _inv(num_points, idx, rs, attrs):
    r = _o.invert_neighbors_list(num_points, idx, rs, attrs)
    return r.neighbors_index, r.neighbors_row_splits, r.neighbors_attributes


_lib.impl("invert_neighbors_list", _inv, "CPU")
_lib.impl("reduce_subarrays_sum", _o.reduce_subarrays_sum, "CPU")
_lib.impl("sparse_conv", lambda *a: _o.sparse_conv(*a), "CPU")
_lib.impl("continuous_conv", lambda *a: _o.continuous_conv(*a[:10], normalize=a[10]), "CPU")


class InvertNeighborsListResult(NamedTuple):
    neighbors_index: torch.Tensor
    neighbors_row_splits: torch.Tensor
    neighbors_attributes: torch.Tensor


def invert_neighbors_list(num_points: int, inp_neighbors_index: torch.Tensor, inp_neighbors_row_splits: torch.Tensor,
                          inp_neighbors_attributes: torch.Tensor):
    a, b, c = torch.ops.open3d_oracle.invert_neighbors_list(num_points, inp_neighbors_index,
                                                            inp_neighbors_row_splits, inp_neighbors_attributes)
    return InvertNeighborsListResult(a, b, c)


def reduce_subarrays_sum(values, row_splits):
    return torch.ops.open3d_oracle.reduce_subarrays_sum(values, row_splits)


def sparse_conv(filters, inp_features, inp_importance, neighbors_index, neighbors_kernel_index,
                neighbors_importance, neighbors_row_splits, normalize=False, max_temp_mem_MB=64):
    return torch.ops.open3d_oracle.sparse_conv(filters, inp_features, inp_importance, neighbors_index,
                                               neighbors_kernel_index, neighbors_importance,
                                               neighbors_row_splits, normalize)


def continuous_conv(filters, out_positions, extents, offset, inp_positions, inp_features, inp_importance,
                    neighbors_index, neighbors_importance, neighbors_row_splits, align_corners=False,
                    coordinate_mapping="ball_to_cube_radial", normalize=False, interpolation="linear",
                    max_temp_mem_MB=64):
    if not (align_corners and coordinate_mapping == "ball_to_cube_radial" and interpolation == "linear"):
        raise NotImplementedError("oracle restates only the configuration the reference uses")
    return torch.ops.open3d_oracle.continuous_conv(filters, out_positions, extents, offset, inp_positions,
                                                   inp_features, inp_importance, neighbors_index,
                                                   neighbors_importance, neighbors_row_splits, normalize)
