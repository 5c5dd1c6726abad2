"""Drop-in for the reference's python module `adaptivesurfacereconstruction`
(pybind11 module asrpybind, cpp/pybind/module.cpp:279-489, re-exported by
python/adaptivesurfacereconstruction/__init__.py:16-18) on the asr_b200 CUDA
backend.  Same function names, keyword names, defaults, array dtypes/shapes and
exception types; arrays go in and out as numpy (host) buffers exactly like the
reference, the work happens on the current CUDA device.

In scope (hot path): create_octree, create_grids_from_octree,
create_dual_vertex_indices, reconstruct_surface (octree-conv -> SDF -> vertices).
The "next" rows f-1 (triangle connectivity), f-2 (KDTree: radius estimation, outlier and
density pre-filters) and f-3 (remove_connected_components) run on the GPU as well.
"""
import os
import warnings

import numpy as np
import torch

from asr_b200 import model as _model
from asr_b200 import ops as _ops
from asr_b200 import pipeline as _pipeline

__version__ = "0.2.0"
INT64_MAX = np.iinfo(np.int64).max


def get_version_str():
    """module.cpp:284 — the reference returns ASR_VERSION."""
    return __version__ + "+b200"


def get_third_party_notices():
    """module.cpp:287."""
    return "asr_b200 uses PyTorch (BSD-3-Clause) and NVIDIA CUB (Apache-2.0 with LLVM exception)."


class Octree:
    """Opaque handle, no python methods (module.cpp:282)."""

    def __init__(self, impl):
        self._impl = impl


def _f32(a, name, shape_msg, ndim, last=None):
    a = np.ascontiguousarray(a, dtype=np.float32)  # py::array::c_style | forcecast
    if a.ndim != ndim or (last is not None and a.shape[-1] != last):
        raise ValueError(shape_msg)
    return a


def create_octree(points, radii, bb_min, bb_max, radius_scale=1, grow_steps=0, max_depth=21):
    """module.cpp:372-374 / pyCreateOctreeFromPoints :144-161."""
    points = _f32(points, "points", "points must have shape [N,3]", 2, 3)
    radii = np.ascontiguousarray(radii, dtype=np.float32)
    if radii.ndim != 1 or radii.shape[0] != points.shape[0]:
        raise ValueError("radii must have shape [N]")
    t = _ops.Octree(torch.from_numpy(points).cuda(), torch.from_numpy(radii).cuda(), bb_min, bb_max,
                    float(radius_scale), int(grow_steps), int(max_depth))
    return Octree(t)


def create_grids_from_octree(tree, num_levels, voxel_info_all_levels=False):
    """module.cpp:402-403 / pyCreateGridsFromOctree :163-228: list (finest first)
    of dicts of numpy arrays; a key is omitted when its array is empty."""
    grids = tree._impl.grids(int(num_levels), bool(voxel_info_all_levels))
    out = []
    for g in grids:
        d = {}
        for k, v in g.items():
            a = v.cpu().numpy()
            d[k] = a.view(np.uint64) if k == "voxel_keys" else a
        out.append(d)
    return out


def create_dual_vertex_indices(tree):
    """module.cpp:443 / :230-235: size_t[num_duals, 8]."""
    return tree._impl.dual_vertex_indices().cpu().numpy().view(np.uint64)


_MODEL_CACHE = {}


def _load_model(levels=5):
    """The reference loads `model.pt` from the resource dir (asr.cpp:50,138-141:
    env ASR_RESOURCE_DIR or <module dir>/asr_resources).  Accepts a TorchScript
    archive or a plain state_dict file with the reference's key names."""
    res = os.environ.get("ASR_RESOURCE_DIR", os.path.join(os.path.dirname(__file__), "asr_resources"))
    path = os.path.join(res, "model.pt")
    if path in _MODEL_CACHE:
        return _MODEL_CACHE[path]
    if not os.path.exists(path):
        raise RuntimeError("model.pt not found in the resource dir %r (set ASR_RESOURCE_DIR); the released weights "
                           "are not redistributable offline" % res)
    net = _model.from_state_dict(_model.load_weights_file(path), levels)
    _MODEL_CACHE[path] = net
    return net


def reconstruct_surface(points, normals, radii, point_radius_scale=1.0, density_percentile_threshold=10.0,
                        point_radius_estimation_knn=24, octree_max_depth=21, contouring_value_threshold=1.0,
                        keep_n_connected_components=INT64_MAX, minimum_component_size=3, model=None):
    """module.cpp:291-306 / pyReconstructSurface :58-109 -> asr::ReconstructSurface
    (asr.cpp:95-349).  Runs the hot path (grid building, aggregation search,
    aggregate/unet/decode, dual-contouring vertices) on the GPU.

    Pre-filter as in asr.cpp:114-138: without radii they are estimated from the
    `point_radius_estimation_knn` nearest neighbours and outliers are dropped; with radii the
    sparsest `density_percentile_threshold` percent of the points are dropped (see
    asr_b200.ops.density_inlier for the one deliberate difference to the reference).
    `model` (an asr_b200.model.UNet) overrides the model.pt lookup.
    Limit of this backend: point_radius_estimation_knn <= 64 (ValueError above; the reference accepts any k)."""
    points = _f32(points, "points", "points must have shape [num_points,3]", 2, 3)
    normals = np.ascontiguousarray(normals, dtype=np.float32)
    if normals.ndim != 2 or normals.shape != points.shape:
        raise ValueError("normals must have shape [num_points,3]")
    radii = np.ascontiguousarray(radii if radii is not None else [], dtype=np.float32)
    if radii.size > 0 and (radii.ndim != 1 or radii.shape[0] != points.shape[0]):
        raise ValueError("radii must have shape [num_point3]")
    if points.shape[0] == 0:
        raise RuntimeError("points is null!\n")
    # preprocessing (asr.cpp:114-138): estimate the radii from the k nearest neighbours and drop
    # outliers, or, with given radii, drop the sparsest points
    dev = torch.device("cuda")
    p = torch.from_numpy(points).to(dev)
    tree = _ops.KDTree(p)
    if radii.size == 0:
        r = tree.compute_k_radius(int(point_radius_estimation_knn))
        keep = tree.compute_inlier(r, 0.5, int(point_radius_estimation_knn), 1)
    else:
        r = torch.from_numpy(radii).to(dev)
        keep = _ops.density_inlier(tree.compute_radius_neighbors(r), float(density_percentile_threshold))
    del tree
    p, r = p[keep].contiguous(), r[keep].contiguous()
    nrm = torch.from_numpy(normals).to(dev)[keep].contiguous()
    if p.shape[0] == 0:
        raise RuntimeError("points is null!\n")
    net = model if model is not None else _load_model()
    out = _pipeline.reconstruct_vertices(net, p, nrm, r, p.min(0).values.cpu().numpy(), p.max(0).values.cpu().numpy(),
                                         radius_scale=float(point_radius_scale), max_depth=int(octree_max_depth),
                                         contouring_value_threshold=float(contouring_value_threshold), triangles=True)
    # asr.cpp:343-345: RemoveConnectedComponents with the caller's limits
    v, t = _ops.remove_connected_components(out["vertices"], out["triangles"], int(keep_n_connected_components),
                                            int(minimum_component_size))
    return {"vertices": v.cpu().numpy(), "triangles": t.cpu().numpy()}


def remove_connected_components(vertices, triangles, keep_n_largest_components, minimum_component_size=3):
    """module.cpp:348-350 / pyRemoveConnectedComponents :111-142 (row f-3)."""
    vertices = _f32(vertices, "vertices", "vertices must have shape [N,3]", 2, 3)
    triangles = np.ascontiguousarray(triangles, dtype=np.int32)
    if triangles.ndim != 2 or triangles.shape[1] != 3:
        raise ValueError("triangles must have shape [N,3]")
    if triangles.size and (triangles.min() < 0 or triangles.max() >= vertices.shape[0]):
        raise RuntimeError("triangle index out of range")  # std::out_of_range in the reference
    v, t = _ops.remove_connected_components(torch.from_numpy(vertices).cuda(), torch.from_numpy(triangles).cuda(),
                                            int(keep_n_largest_components), int(minimum_component_size))
    return {"vertices": v.cpu().numpy(), "triangles": t.cpu().numpy()}


class KDTree:
    """module.cpp:455-489 (row f-2): nearest-neighbour statistics of a cloud, on the GPU."""

    def __init__(self, points):
        points = np.ascontiguousarray(points, dtype=np.float32)
        if points.ndim != 2 or points.shape[1] != 3:
            raise ValueError("points must have shape [N,3]")
        self._num = points.shape[0]
        self._impl = _ops.KDTree(torch.from_numpy(points).cuda())

    def compute_k_radius(self, k):
        """module.cpp:459 / pyComputeKRadius: distance to the k-th nearest point (itself included)."""
        return self._impl.compute_k_radius(int(k)).cpu().numpy()

    def _radii(self, radii):
        radii = np.ascontiguousarray(radii, dtype=np.float32)
        if radii.ndim != 1 or radii.shape[0] != self._num:
            raise ValueError("radii must have shape [N]")
        return torch.from_numpy(radii).cuda()

    def compute_inlier(self, radii, radius_fraction=0.5, k=24, outlier_threshold=1):
        """module.cpp:468."""
        return self._impl.compute_inlier(self._radii(radii), float(radius_fraction), int(k),
                                         int(outlier_threshold)).cpu().numpy()

    def compute_radius_neighbors(self, radii):
        """module.cpp:483: number of points within each point's radius, as a list of ints."""
        return self._impl.compute_radius_neighbors(self._radii(radii)).cpu().tolist()
