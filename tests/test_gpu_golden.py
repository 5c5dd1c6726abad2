"""GPU: the CUDA path reproduces the reference-generated golden fixture
(tests/golden/ref_sphere3k.npz): index arrays bit-exact, floats within 1e-4."""
import os

import numpy as np
import pytest
import torch

from helpers import dev

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_sphere3k.npz")


def test_cuda_path_reproduces_golden():
    from asr_b200 import model, pipeline
    from oracle import model_cpu
    g = np.load(GOLD)
    P = model_cpu.init_params(5, seed=int(g["weights_seed"]), stress=True)
    net = model.from_state_dict(P, 5)
    out = pipeline.reconstruct_vertices(net, dev(g["points"]), dev(g["normals"]), dev(g["radii"]), g["bb_min"],
                                        g["bb_max"])
    d = out["input_dict"]
    assert np.array_equal(out["octree"].leaves().cpu().numpy().view(np.uint64), g["leaves"])
    for name in g.files:
        if name.startswith("grid") and not name.endswith("voxel_keys"):
            assert np.array_equal(d[name[6:] + name[4]].cpu().numpy(), g[name]), name
    for k in ("aggregation_neighbors_index", "aggregation_neighbors_dist", "aggregation_row_splits"):
        assert np.array_equal(d[k].cpu().numpy(), g[k]), k
    assert np.abs(d["aggregation_scale_compat"].cpu().numpy() - g["aggregation_scale_compat"]).max() <= 1e-6
    assert np.array_equal(out["dual_vertex_indices"].cpu().numpy().astype(np.uint64), g["dual_vertex_indices"])
    feats, imp = net.aggregate(d)
    assert np.abs(feats.cpu().numpy() - g["aggregate_feats"]).max() <= 1e-4
    assert np.abs(imp.cpu().numpy() - g["aggregate_importance"]).max() <= 1e-6
    code = net.unet((feats, imp), d)
    assert np.abs(code.cpu().numpy() - g["code"]).max() <= 1e-4
    assert np.abs(out["values"].cpu().numpy() - g["values"]).max() <= 1e-4
    v, grad = net.decode_with_gradient(torch.full((code.shape[0], 3), 0.25, device="cuda"), code)
    assert np.abs(grad.cpu().numpy() - g["decode_grad"]).max() <= 1e-4
    # contouring of the golden values: same intersecting duals, same vertices as the reference mesh
    from asr_b200 import ops
    verts, vd = ops.contour_vertices(dev(g["values"]), out["dual_vertex_indices"], d["voxel_centers0"], 1.0)
    assert verts.shape[0] > 100
    assert np.array_equal(verts.cpu().numpy(), g["mesh_vertices"][:verts.shape[0]])
