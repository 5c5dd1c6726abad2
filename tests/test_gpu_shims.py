"""GPU: the drop-in boundaries called the way the reference calls them
(cpp/pybind/module.cpp signatures; models/common_torch.py:124-142 and
models/v0/net_definitions_torch.py:22-36,59-70,108-116 call patterns)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import dev

pytestmark = pytest.mark.gpu


def test_python_module_drop_in_matches_oracle(geom_checkers):
    import adaptivesurfacereconstruction as asr
    from asr_b200 import clouds
    c = clouds.adaptive_blob(15000, seed=5)
    tree = asr.create_octree(c["points"], c["radii"], c["bb_min"], c["bb_max"], radius_scale=1, grow_steps=0,
                             max_depth=21)
    grids = asr.create_grids_from_octree(tree, 5, voxel_info_all_levels=True)
    duals = asr.create_dual_vertex_indices(tree)
    cname, Cls = geom_checkers[-1]
    o = Cls(c["points"], c["radii"], c["bb_min"], c["bb_max"])
    og = o.grids(5, True)
    assert len(grids) == 5
    for g, r in zip(grids, og):
        assert set(g) == set(r)
        for k, v in r.items():
            assert isinstance(g[k], np.ndarray) and g[k].dtype == v.dtype and np.array_equal(g[k], v), k
    assert duals.dtype == np.uint64 and np.array_equal(duals, o.dual_vertex_indices())
    g1 = asr.create_grids_from_octree(tree, 2)  # voxel_info_all_levels defaults to False
    assert "voxel_centers" in g1[0] and "voxel_centers" not in g1[1] and "up_neighbors_index" not in g1[1]


def test_reconstruct_surface_drop_in():
    import adaptivesurfacereconstruction as asr
    from asr_b200 import clouds, model
    c = clouds.sphere(6000, seed=2)
    net = model.seeded_weights(model.UNet(5), seed=4).cuda()
    mesh = asr.reconstruct_surface(c["points"], c["normals"], c["radii"], model=net, contouring_value_threshold=1e9)
    assert set(mesh) == {"vertices", "triangles"}
    assert mesh["vertices"].dtype == np.float32 and mesh["vertices"].shape[1] == 3
    assert mesh["triangles"].dtype == np.int32 and mesh["triangles"].shape[1] == 3
    # random weights give an arbitrary SDF: only the structure is checked here (the triangle and
    # component passes are compared with the reference in test_gpu_geometry.py).  Every vertex that
    # survives the default component filter (>= 3 vertices) is used by a triangle.
    assert len(np.unique(mesh["triangles"])) == mesh["vertices"].shape[0]
    if mesh["triangles"].size:
        assert mesh["triangles"].min() >= 0 and mesh["triangles"].max() < mesh["vertices"].shape[0]
    out = asr.remove_connected_components(mesh["vertices"], mesh["triangles"], 1)
    assert out["vertices"].shape[0] <= mesh["vertices"].shape[0] and out["triangles"].shape[1] == 3
    with pytest.raises(RuntimeError):
        asr.reconstruct_surface(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), np.zeros(0, np.float32),
                                model=net)


def test_open3d_ops_called_like_the_reference():
    import open3d.ml.torch as ml3d
    from open3d.ml.torch import ops
    from asr_b200 import clouds, ops as k
    from oracle import ops_cpu
    c = clouds.adaptive_blob(12000, seed=6)
    t = k.Octree(dev(c["points"]), dev(c["radii"]), c["bb_min"], c["bb_max"])
    grids = t.grids(2, True)
    g = grids[0]
    V = g["neighbors_row_splits"].shape[0] - 1
    gen = torch.Generator().manual_seed(0)
    kernel = (torch.rand((55, 32, 56), generator=gen) - 0.5) * 0.2
    x = torch.rand((V, 32), generator=gen)
    imp = torch.rand(V, generator=gen)
    # SpecialSparseConv.forward, common_torch.py:124-142 (note the CPU-constructed empty tensors)
    neighbors_importance = imp.cuda()[g["neighbors_index"].to(torch.int64)]
    out_importance = ops.reduce_subarrays_sum(neighbors_importance, g["neighbors_row_splits"])
    out = ops.sparse_conv(filters=kernel.cuda(), inp_features=x.cuda(),
                          inp_importance=torch.empty((0,), dtype=torch.float32),
                          neighbors_index=g["neighbors_index"], neighbors_kernel_index=g["neighbors_kernel_index"],
                          neighbors_importance=neighbors_importance, neighbors_row_splits=g["neighbors_row_splits"],
                          normalize=True)
    cpu = {kk: v.cpu() for kk, v in g.items()}
    nimp = imp[cpu["neighbors_index"].long()]
    ref = ops_cpu.sparse_conv(kernel, x, torch.empty(0), cpu["neighbors_index"], cpu["neighbors_kernel_index"], nimp,
                              cpu["neighbors_row_splits"], True, dtype=torch.float64)
    assert (out.cpu().double() - ref).abs().max() <= 1e-4
    assert (out_importance.cpu() - ops_cpu.reduce_subarrays_sum(nimp, cpu["neighbors_row_splits"])).abs().max() <= 1e-5
    out2 = ops.sparse_conv(filters=kernel.cuda(), inp_features=x.cuda(),
                           inp_importance=torch.empty((0,), dtype=torch.float32),
                           neighbors_index=g["neighbors_index"], neighbors_kernel_index=g["neighbors_kernel_index"],
                           neighbors_importance=torch.empty((0,), dtype=torch.float32),
                           neighbors_row_splits=g["neighbors_row_splits"], normalize=False)
    ref2 = ops_cpu.sparse_conv(kernel, x, torch.empty(0), cpu["neighbors_index"], cpu["neighbors_kernel_index"],
                               torch.empty(0), cpu["neighbors_row_splits"], False, dtype=torch.float64)
    assert (out2.cpu().double() - ref2).abs().max() <= 1e-4

    # invert_neighbors_list inside torch.jit.script, net_definitions_torch.py:22-36
    @torch.jit.script
    def invert_script(num_points_tensor, neighbors_index, neighbors_row_splits, neighbors_kernel_index):
        ans = ml3d.ops.invert_neighbors_list(num_points_tensor.shape[0], neighbors_index, neighbors_row_splits,
                                             neighbors_kernel_index)
        return ans

    ans = invert_script(grids[1]["voxel_centers"], g["up_neighbors_index"], g["up_neighbors_row_splits"],
                        g["up_neighbors_kernel_index"])
    r = ops_cpu.invert_neighbors_list(grids[1]["voxel_centers"].shape[0], cpu["up_neighbors_index"],
                                      cpu["up_neighbors_row_splits"], cpu["up_neighbors_kernel_index"])
    assert torch.equal(ans.neighbors_index.cpu(), r.neighbors_index)
    assert torch.equal(ans.neighbors_row_splits.cpu(), r.neighbors_row_splits)
    assert torch.equal(ans.neighbors_attributes.cpu(), r.neighbors_attributes)

    # ContinuousConv layer, net_definitions_torch.py:59-70,108-116
    conv = ml3d.layers.ContinuousConv(in_channels=4, filters=32, activation=F.relu, kernel_size=[4, 4, 4],
                                      coordinate_mapping="ball_to_cube_radial", normalize=True).cuda()
    assert set(dict(conv.named_parameters())) == {"kernel", "bias", "offset"}
    assert tuple(conv.kernel.shape) == (4, 4, 4, 4, 32)
    pts = dev(c["points"])
    idx, d2, rs = k.multi_radius_search(pts, g["voxel_centers"], g["voxel_sizes"])
    nimp = torch.rand(idx.shape[0], generator=gen).cuda()
    feats = torch.rand((pts.shape[0], 4), generator=gen).cuda()
    with torch.no_grad():
        y = conv(feats, pts, g["voxel_centers"], extents=g["voxel_sizes"], user_neighbors_index=idx,
                 user_neighbors_row_splits=rs, user_neighbors_importance=nimp)
        ref = ops_cpu.continuous_conv(conv.kernel.cpu(), g["voxel_centers"].cpu(), g["voxel_sizes"].cpu(),
                                      torch.zeros(3), pts.cpu(), feats.cpu(), torch.empty(0), idx.cpu(), nimp.cpu(),
                                      rs.cpu(), normalize=True, dtype=torch.float64)
    assert (y.cpu().double() - torch.relu(ref + conv.bias.cpu().double())).abs().max() <= 1e-4


def test_asrtool_ply_to_ply(tmp_path):
    """Row f-4: the command line tool end to end (seeded random weights: model.pt is not available offline)."""
    from asr_b200 import asrtool, clouds, plyio
    c = clouds.sphere(5000, seed=3)
    src = tmp_path / "cloud.ply"
    with open(src, "wb") as f:
        f.write(("ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % len(c["points"]) + "".join(
            "property float %s\n" % n for n in ("x", "y", "z", "nx", "ny", "nz", "radius")) + "end_header\n").encode())
        f.write(np.concatenate([c["points"], c["normals"], c["radii"][:, None]], 1).astype("<f4").tobytes())
    dst = tmp_path / "mesh.ply"
    assert asrtool.main(["--in", str(src), "--out", str(dst), "--random-weights", "4"]) == 0
    v, t = plyio.read_mesh(str(dst))
    assert v.shape[1] == 3 and t.shape[1] == 3 and (t.size == 0 or t.max() < len(v))
    assert asrtool.main(["--version"]) == 0 and asrtool.main([]) == 1
