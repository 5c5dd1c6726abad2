"""`open3d.ml.torch.ops` — the four ops the reference calls, registered as
`torch.library` ops in the `open3d::` namespace (scriptable, so the reference's
@torch.jit.script wrapper net_definitions_torch.py:22-36 and torch.jit.trace_module
keep working) and implemented by libasr_b200.so.  CUDA tensors only: a CPU
tensor raises, there is no fallback."""
from typing import NamedTuple

import torch

from asr_b200 import ops as _k

# Schemas = Open3D v0.14.1's own registrations (cpp/open3d/ml/pytorch/**/*Ops.cpp), INCLUDING the
# defaulted trailing arguments, so that a TorchScript archive traced against real Open3D
# (models/v0/convert_tf2torchscript.py:117-122 -> model.pt, loaded by asr.cpp:138-141) resolves its
# `open3d::*` calls against this library.
_lib = torch.library.Library("open3d", "DEF")
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

_PLANS = {}


def _plan(idx, kidx, rs, K):
    """Plans are cached per neighbour table (data pointers + sizes + version)."""
    key = (idx.data_ptr(), kidx.data_ptr(), rs.data_ptr(), idx.numel(), rs.numel(), K, idx._version, kidx._version)
    p = _PLANS.get(key)
    if p is None:
        if len(_PLANS) > 64:
            _PLANS.clear()
        p = (_k.ConvPlan(idx, kidx, rs, K), idx, kidx, rs)  # keep the tensors alive with the plan
        _PLANS[key] = p
    return p[0]


def _cuda_invert(num_points, idx, rs, attrs):
    r = _k.invert_neighbors_list(num_points, idx, rs, attrs)
    return r.neighbors_index, r.neighbors_row_splits, r.neighbors_attributes


def _cuda_reduce(values, row_splits):
    return _k.reduce_subarrays_sum(values, row_splits)


def _on(t, like):
    """The reference passes CPU-constructed empty tensors for unused inputs (common_torch.py:130-136)."""
    return t.to(like.device) if t.device != like.device and t.numel() == 0 else t


def _cuda_sparse_conv(filters, x, inp_importance, idx, kidx, nimp, rs, normalize=False, max_temp_mem_MB=64):
    # max_temp_mem_MB bounds Open3D's temporary B-matrix; this implementation has no such buffer
    inp_importance, nimp = _on(inp_importance, x), _on(nimp, x)
    plan = _plan(idx, kidx, rs, filters.shape[0])
    has_imp = nimp.numel() > 0 or inp_importance.numel() > 0
    norm = None
    if normalize and nimp.numel() > 0:
        norm = _k.reduce_subarrays_sum(nimp, rs)
    cout = filters.shape[2]
    if filters.shape[1] % 4 or cout % 4:
        raise ValueError("asr_b200 sparse_conv needs channel counts that are multiples of 4")
    return _k.sparse_conv(plan, filters, x, inp_importance=inp_importance if inp_importance.numel() else None,
                          neighbors_importance=nimp if nimp.numel() else None, importance_col=0 if has_imp else cout,
                          normalize=normalize, normalize_col=0, normalizer=norm)


def _cuda_cconv(filters, out_positions, extents, offset, inp_positions, inp_features, inp_importance, idx, nimp, rs,
                align_corners=False, coordinate_mapping="ball_to_cube_radial", normalize=False, interpolation="linear",
                max_temp_mem_MB=64):
    if not (align_corners and coordinate_mapping == "ball_to_cube_radial" and interpolation == "linear"):
        raise NotImplementedError("asr_b200 implements the configuration the reference uses: align_corners=True, "
                                  "coordinate_mapping='ball_to_cube_radial', interpolation='linear'")
    inp_importance, nimp = _on(inp_importance, inp_features), _on(nimp, inp_features)
    return _k.continuous_conv(filters, out_positions, extents, offset, inp_positions, inp_features,
                              inp_importance if inp_importance.numel() else None, idx,
                              nimp if nimp.numel() else None, rs, normalize=normalize)


_lib.impl("invert_neighbors_list", _cuda_invert, "CUDA")
_lib.impl("reduce_subarrays_sum", _cuda_reduce, "CUDA")
_lib.impl("sparse_conv", _cuda_sparse_conv, "CUDA")
_lib.impl("continuous_conv", _cuda_cconv, "CUDA")


def _no_cpu(*a, **k):
    raise RuntimeError("asr_b200: open3d:: ops run on CUDA tensors only (no CPU fallback); move the inputs to the GPU")


for _n in ("invert_neighbors_list", "reduce_subarrays_sum", "sparse_conv", "continuous_conv"):
    _lib.impl(_n, _no_cpu, "CPU")


class InvertNeighborsListResult(NamedTuple):
    neighbors_index: torch.Tensor
    neighbors_row_splits: torch.Tensor
    neighbors_attributes: torch.Tensor


def invert_neighbors_list(num_points: int, inp_neighbors_index: torch.Tensor, inp_neighbors_row_splits: torch.Tensor,
                          inp_neighbors_attributes: torch.Tensor):
    a, b, c = torch.ops.open3d.invert_neighbors_list(num_points, inp_neighbors_index, inp_neighbors_row_splits,
                                                     inp_neighbors_attributes)
    return InvertNeighborsListResult(a, b, c)


def reduce_subarrays_sum(values, row_splits):
    return torch.ops.open3d.reduce_subarrays_sum(values, row_splits)


def sparse_conv(filters, inp_features, inp_importance, neighbors_index, neighbors_kernel_index,
                neighbors_importance, neighbors_row_splits, normalize=False, max_temp_mem_MB=64):
    return torch.ops.open3d.sparse_conv(filters, inp_features, inp_importance, neighbors_index,
                                        neighbors_kernel_index, neighbors_importance, neighbors_row_splits,
                                        normalize, max_temp_mem_MB)


def continuous_conv(filters, out_positions, extents, offset, inp_positions, inp_features, inp_importance,
                    neighbors_index, neighbors_importance, neighbors_row_splits, align_corners=False,
                    coordinate_mapping="ball_to_cube_radial", normalize=False, interpolation="linear",
                    max_temp_mem_MB=64):
    return torch.ops.open3d.continuous_conv(filters, out_positions, extents, offset, inp_positions, inp_features,
                                            inp_importance, neighbors_index, neighbors_importance,
                                            neighbors_row_splits, align_corners, coordinate_mapping, normalize,
                                            interpolation, max_temp_mem_MB)
