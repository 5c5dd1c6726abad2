"""TEST INFRASTRUCTURE ONLY — the whole hot path on the CPU, sequenced like
asr::ReconstructSurface (reference cpp/lib/asr.cpp:143-342): geometry from the
compiled reference TUs (oracle/_ref) when present, else from the port
(oracle/geom_oracle.cpp); search / network from oracle/ops_cpu.py and
oracle/model_cpu.py.  Used as the parity checker of the GPU pipeline and as the
`cpu_baseline` / `--impl reference` leg of bench.py.  Never used by the product.
"""
import time

import numpy as np
import torch

from . import geomlib, model_cpu, ops_cpu, reflib


def geometry_backend(prefer_reference=True):
    if prefer_reference and reflib.available():
        return "reference", reflib.RefOctree
    return "port", geomlib.PortOctree


def build_input_dict(cloud, levels=5, radius_scale=1.0, max_depth=21, prefer_reference=True, times=None):
    kind, Cls = geometry_backend(prefer_reference)
    t0 = time.perf_counter()
    tree = Cls(cloud["points"], cloud["radii"], cloud["bb_min"], cloud["bb_max"], radius_scale, 0, max_depth)
    t1 = time.perf_counter()
    duals = tree.dual_vertex_indices()
    t2 = time.perf_counter()
    grids = tree.grids(levels, True)
    t3 = time.perf_counter()
    n = cloud["points"].shape[0]
    d = {"points": torch.from_numpy(cloud["points"]),
         "feats": torch.from_numpy(np.concatenate([cloud["normals"], np.ones((n, 1), np.float32)], 1))}
    for i, g in enumerate(grids):
        for k, v in g.items():
            if k != "voxel_keys":
                d[k + str(i)] = torch.from_numpy(v)
    idx, d2, rs = ops_cpu.multi_radius_search(cloud["points"], grids[0]["voxel_centers"], grids[0]["voxel_sizes"])
    sc = ops_cpu.scale_compatibility(grids[0]["voxel_sizes"], cloud["radii"], idx, rs)
    t4 = time.perf_counter()
    d["aggregation_neighbors_index"] = torch.from_numpy(idx)
    d["aggregation_neighbors_dist"] = torch.from_numpy(d2)
    d["aggregation_row_splits"] = torch.from_numpy(rs)
    d["aggregation_scale_compat"] = torch.from_numpy(sc)
    if times is not None:
        times.update({"octree": t1 - t0, "duals": t2 - t1, "grids": t3 - t2, "search": t4 - t3, "geometry": kind})
    return d, duals


def run(cloud, params, levels=5, threshold=1.0, dtype=None, prefer_reference=True, times=None):
    """Returns dict(values [V0,2], vertices [M,3], vertex_dual [M], input_dict, duals)."""
    times = times if times is not None else {}
    d, duals = build_input_dict(cloud, levels, prefer_reference=prefer_reference, times=times)
    t0 = time.perf_counter()
    with torch.no_grad():
        feats = model_cpu.aggregate(params, d, dtype=dtype)
        t1 = time.perf_counter()
        code = model_cpu.unet(params, feats, d, levels, dtype=dtype)
        t2 = time.perf_counter()
        values = model_cpu.decode(params, torch.zeros(code.shape[0], 3), code)
        values = values.clone()
        values[:, 0] *= d["voxel_sizes0"].to(values.dtype)
    t3 = time.perf_counter()
    v32 = values.to(torch.float32).numpy()
    verts, vdual = geomlib.contour_vertices(v32, duals, d["voxel_centers0"].numpy(), threshold)
    t4 = time.perf_counter()
    times.update({"aggregate": t1 - t0, "unet": t2 - t1, "decode": t3 - t2, "contour": t4 - t3})
    return {"values": values, "vertices": verts, "vertex_dual": vdual, "input_dict": d, "duals": duals}
