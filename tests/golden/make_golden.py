"""Generates tests/golden/ref_sphere3k.npz — run ONCE in the build container,
where /root/reference exists:   python tests/golden/make_golden.py

What is "golden" here (SURVEY.md §8c — the reference ships no test vectors):
  * every integer / geometry array comes from the reference's OWN code: the
    unmodified cpp/lib translation units compiled into oracle/_ref
    (CreateOctreeFromPoints, CreateGridsFromOctree, CreateDualVertexIndices,
    CreateTriangleMesh);
  * the network outputs come from the reference's OWN model definition
    (models/v0/net_definitions_torch.py, imported from /root/reference) driven
    exactly like asr.cpp:315-336, with the Open3D ops (absent in this image)
    supplied by oracle/o3d_shim -> oracle/ops_cpu.py.  So the *topology* is the
    reference's; the op arithmetic is the restatement ("parity unpinned").
Weights: oracle.model_cpu.init_params(5, seed=11, stress=True) (torch CPU RNG).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "o3d_shim"))
sys.path.insert(0, os.path.join(ROOT, "adaptive-surface-reconstruction_b200", "asr_b200"))
sys.path.insert(0, "/root/reference")

import clouds  # noqa: E402  (pure-numpy generator module, imported without the CUDA package)
from models.v0.net_definitions_torch import UNet5  # noqa: E402  (the reference's model)
from oracle import model_cpu, ops_cpu, reflib  # noqa: E402

SEED_W = 11


def main():
    c = clouds.sphere(3000, seed=5)
    c["radii"] = (c["radii"] * np.random.default_rng(6).uniform(0.6, 2.5, 3000)).astype(np.float32)
    tree = reflib.RefOctree(c["points"], c["radii"], c["bb_min"], c["bb_max"])
    grids = tree.grids(5, True)
    duals = tree.dual_vertex_indices()
    out = {"points": c["points"], "normals": c["normals"], "radii": c["radii"], "bb_min": c["bb_min"],
           "bb_max": c["bb_max"], "leaves": tree.leaves(), "dual_vertex_indices": duals}
    inp = {"points": torch.from_numpy(c["points"]),
           "feats": torch.from_numpy(np.concatenate([c["normals"], np.ones((3000, 1), np.float32)], 1))}
    for i, g in enumerate(grids):
        for k, v in g.items():
            out["grid%d_%s" % (i, k)] = v
            if k != "voxel_keys":
                inp[k + str(i)] = torch.from_numpy(v)
    idx, d2, rs = ops_cpu.multi_radius_search(c["points"], grids[0]["voxel_centers"], grids[0]["voxel_sizes"])
    sc = ops_cpu.scale_compatibility(grids[0]["voxel_sizes"], c["radii"], idx, rs)
    inp["aggregation_neighbors_index"] = torch.from_numpy(idx)
    inp["aggregation_neighbors_dist"] = torch.from_numpy(d2)
    inp["aggregation_row_splits"] = torch.from_numpy(rs)
    inp["aggregation_scale_compat"] = torch.from_numpy(sc)
    out.update({"aggregation_neighbors_index": idx, "aggregation_neighbors_dist": d2, "aggregation_row_splits": rs,
                "aggregation_scale_compat": sc})

    net = UNet5(with_importance="all", normalized_channels=8, residual_skip_connection=True)
    P = model_cpu.init_params(5, seed=SEED_W, stress=True)
    net.load_state_dict(P)
    with torch.no_grad():
        feats = net.aggregate(inp)
        code = net.unet(feats, inp)
        values = net.decode(torch.zeros(code.shape[0], 3), code).contiguous()
        vg, grad = net.decode_with_gradient(torch.full((code.shape[0], 3), 0.25), code)
    vals = values.numpy().copy()
    vals[:, 0] *= grids[0]["voxel_sizes"]  # asr.cpp:334-336
    mesh = reflib.create_triangle_mesh(vals, duals, grids[0]["voxel_centers"], 1.0)
    out.update({"aggregate_feats": feats[0].numpy(), "aggregate_importance": feats[1].numpy(), "code": code.numpy(),
                "values": vals, "decode_grad_values": vg.numpy(), "decode_grad": grad.numpy(),
                "mesh_vertices": mesh["vertices"], "mesh_triangles": mesh["triangles"],
                "weights_seed": np.int64(SEED_W), "torch_version": np.array(torch.__version__)})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_sphere3k.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB; V0 =", len(out["leaves"]), "pairs =", len(idx),
          "duals =", len(duals), "mesh verts/tris =", mesh["vertices"].shape[0], mesh["triangles"].shape[0],
          "|values| max", float(np.abs(vals).max()))


if __name__ == "__main__":
    main()
