"""GPU parity (bit-exact) of the octree / grid / dual-cell / contouring kernels
against the geometry oracles, through the C ABI."""
import numpy as np
import pytest
import torch

from helpers import dev, small_clouds, u64

pytestmark = pytest.mark.gpu


def _build(c, **kw):
    from asr_b200 import ops
    return ops.Octree(dev(c["points"]), dev(c["radii"]), c["bb_min"], c["bb_max"], **kw)


@pytest.mark.parametrize("name,cloud,levels", list(small_clouds()), ids=lambda x: x if isinstance(x, str) else None)
def test_octree_grids_duals_bit_exact(name, cloud, levels, geom_checkers):
    t = _build(cloud)
    leaves = u64(t.leaves())
    grids = t.grids(levels, True)
    duals = t.dual_vertex_indices().cpu().numpy()
    for cname, Cls in geom_checkers:
        o = Cls(cloud["points"], cloud["radii"], cloud["bb_min"], cloud["bb_max"])
        assert np.array_equal(leaves, o.leaves()), cname
        assert t.num_nodes == len(o.nodes()), cname
        for a, b in zip(t.frame(), o.params()):
            assert np.array_equal(a, b), cname
        og = o.grids(levels, True)
        for l, (g, r) in enumerate(zip(grids, og)):
            assert set(g) == set(r), (cname, l, set(g) ^ set(r))
            for k, v in r.items():
                mine = g[k].cpu().numpy()
                if k == "voxel_keys":
                    mine = mine.view(np.uint64)
                assert mine.dtype == v.dtype and mine.shape == v.shape, (cname, l, k, mine.dtype, mine.shape, v.shape)
                assert np.array_equal(mine, v), (cname, l, k)
        od = o.dual_vertex_indices()
        assert np.array_equal(duals.astype(np.uint64), od), cname


def test_grids_without_voxel_info_and_depth_limit(geom_checkers):
    from asr_b200 import clouds
    c = clouds.adaptive_blob(20000, seed=4)
    t = _build(c, max_depth=6)
    grids = t.grids(3, False)
    cname, Cls = geom_checkers[-1]
    o = Cls(c["points"], c["radii"], c["bb_min"], c["bb_max"], max_depth=6)
    og = o.grids(3, False)
    for l, (g, r) in enumerate(zip(grids, og)):
        assert set(g) == set(r), (l, set(g) ^ set(r))
        for k, v in r.items():
            mine = g[k].cpu().numpy()
            if k == "voxel_keys":
                mine = mine.view(np.uint64)
            assert np.array_equal(mine, v), (l, k)


def test_margin_free_bbox_reproduces_junk_leaves(geom_checkers):
    """C++ driver bbox (asr.cpp:148-150): the extreme point maps to INVALID_KEY and
    six junk level-0 leaves 2..7 appear."""
    from asr_b200 import clouds
    c = clouds.sphere(5000, seed=7)
    c["bb_min"], c["bb_max"] = c["points"].min(0), c["points"].max(0)
    t = _build(c)
    leaves = u64(t.leaves())
    cname, Cls = geom_checkers[-1]
    o = Cls(c["points"], c["radii"], c["bb_min"], c["bb_max"])
    assert np.array_equal(leaves, o.leaves())
    assert list(leaves[:6]) == [2, 3, 4, 5, 6, 7]
    g, r = t.grids(5, True), o.grids(5, True)
    for l in range(5):
        for k, v in r[l].items():
            mine = g[l][k].cpu().numpy()
            if k == "voxel_keys":
                mine = mine.view(np.uint64)
            assert np.array_equal(mine, v, equal_nan=True), (l, k)
    assert np.array_equal(t.dual_vertex_indices().cpu().numpy().astype(np.uint64), o.dual_vertex_indices())


def test_empty_and_tiny_inputs():
    from asr_b200 import ops
    z = torch.zeros((0, 3), device="cuda")
    t = ops.Octree(z, torch.zeros(0, device="cuda"), [0, 0, 0], [1, 1, 1])
    assert t.num_leaves == 0
    g = t.grids(2, True)
    assert len(g) == 2 and g[0]["neighbors_row_splits"].tolist() == [0]
    assert t.dual_vertex_indices().shape == (0, 8)
    # a single huge-radius point -> the root is the only leaf
    t = ops.Octree(torch.tensor([[0.5, 0.5, 0.5]], device="cuda"), torch.tensor([10.0], device="cuda"),
                   [0, 0, 0], [1, 1, 1])
    assert u64(t.leaves()).tolist() == [1]
    g = t.grids(2, True)
    assert g[0]["neighbors_index"].tolist() == [0] and g[0]["neighbors_kernel_index"].tolist() == [0]
    # points outside the box are dropped (octree.cpp:248-251)
    t = ops.Octree(torch.tensor([[2.0, 0.5, 0.5]], device="cuda"), torch.tensor([0.1], device="cuda"),
                   [0, 0, 0], [1, 1, 1])
    assert t.num_leaves == 0
    with pytest.raises(ValueError):
        ops.Octree(torch.zeros((4, 2), device="cuda"), torch.zeros(4, device="cuda"), [0, 0, 0], [1, 1, 1])
    with pytest.raises(ValueError):
        ops.Octree(torch.zeros((4, 3), device="cuda"), torch.zeros(3, device="cuda"), [0, 0, 0], [1, 1, 1])


@pytest.mark.parametrize("name,cloud,levels", list(small_clouds())[:3], ids=lambda x: x if isinstance(x, str) else None)
def test_contour_vertices_bit_exact(name, cloud, levels, geom_checkers):
    from asr_b200 import ops
    from oracle import geomlib, reflib
    t = _build(cloud)
    g = t.grids(1, True)[0]
    duals = t.dual_vertex_indices()
    c = g["voxel_centers"].cpu().numpy()
    s = g["voxel_sizes"].cpu().numpy()
    rng = np.random.default_rng(5)
    centre = c.mean(0)
    rad = np.linalg.norm(c - centre, axis=1)
    dist = rad - 0.9 * np.median(rad) + 0.05 * np.sin(9 * c[:, 0])
    vals = np.stack([dist, np.abs(dist) / s * rng.uniform(0.5, 1.5, len(s))], 1).astype(np.float32)
    vals[::97, 0] = 0.0  # exact zeros never count as a sign change
    for thr in (1.0, 0.4):
        v, vd = ops.contour_vertices(dev(vals), duals, g["voxel_centers"], thr)
        pv, pd = geomlib.contour_vertices(vals, duals.cpu().numpy().astype(np.uint64), c, thr)
        assert len(pv) > 20
        assert np.array_equal(vd.cpu().numpy().astype(np.uint64), pd)
        assert np.array_equal(v.cpu().numpy(), pv)
        if reflib.available():
            m = reflib.create_triangle_mesh(vals, duals.cpu().numpy().astype(np.uint64), c, thr)
            assert np.array_equal(m["vertices"][:len(pv)], v.cpu().numpy())


def _norm_tris(t):
    """Rotation-normalised triangle multiset: each triangle starts with its smallest index,
    cyclic order (= orientation) preserved; rows sorted."""
    t = np.asarray(t, np.int64).reshape(-1, 3)
    k = np.argmin(t, 1)
    r = np.stack([t[np.arange(len(t)), (k + i) % 3] for i in range(3)], 1)
    return r[np.lexsort((r[:, 2], r[:, 1], r[:, 0]))]


@pytest.mark.parametrize("name,cloud,levels", list(small_clouds())[:3], ids=lambda x: x if isinstance(x, str) else None)
def test_contour_triangles_match_reference(name, cloud, levels):
    """Triangle part of CreateTriangleMesh (contouring.cpp:202-459) against the compiled
    reference: same triangle multiset up to the rotation of each polygon (the reference starts
    every cycle at a hash-order dependent dual), same fan-centre vertices to float rounding."""
    from asr_b200 import ops
    from oracle import reflib
    if not reflib.available():
        pytest.skip("compiled reference (oracle/_ref) not present")
    t = _build(cloud)
    g = t.grids(1, True)[0]
    duals = t.dual_vertex_indices()
    c = g["voxel_centers"].cpu().numpy()
    s = g["voxel_sizes"].cpu().numpy()
    rng = np.random.default_rng(11)
    centre = c.mean(0)
    rad = np.linalg.norm(c - centre, axis=1)
    dist = rad - 0.9 * np.median(rad) + 0.05 * np.sin(9 * c[:, 0])
    vals = np.stack([dist, np.abs(dist) / s * rng.uniform(0.5, 1.5, len(s))], 1).astype(np.float32)
    for thr in (1.0, 0.4):
        v, tri, vd = ops.contour_mesh(dev(vals), duals, g["voxel_centers"], thr)
        m = reflib.create_triangle_mesh(vals, duals.cpu().numpy().astype(np.uint64), c, thr)
        M = vd.shape[0]
        assert v.shape[0] == m["vertices"].shape[0] and tri.shape[0] == m["triangles"].shape[0] > 50
        assert np.array_equal(v[:M].cpu().numpy(), m["vertices"][:M])
        assert np.abs(v[M:].cpu().numpy() - m["vertices"][M:]).max(initial=0.0) <= 1e-6
        a, b = _norm_tris(tri.cpu().numpy()), _norm_tris(m["triangles"])
        same = (a == b).all(1)
        # a quad whose diagonals are equal to rounding may be split the other way
        assert same.mean() > 0.999, "triangle sets differ: %d of %d" % ((~same).sum(), len(a))


def test_remove_connected_components_matches_reference():
    from asr_b200 import ops
    from oracle import reflib
    if not reflib.available():
        pytest.skip("compiled reference (oracle/_ref) not present")
    rng = np.random.default_rng(3)
    # a handful of separate strips of different sizes + isolated vertices
    verts, tris, base = [], [], 0
    for n in (40, 7, 7, 3, 12, 1, 1, 25):
        verts.append(rng.standard_normal((n, 3)).astype(np.float32))
        for i in range(n - 2):
            tris.append([base + i, base + i + 1, base + i + 2])
        base += n
    verts = np.concatenate(verts)
    tris = np.asarray(tris, np.int32)[rng.permutation(len(tris))]
    for keep, mins in ((np.iinfo(np.int64).max, 3), (3, 3), (2, 1), (4, 8), (1, 100), (100, 1)):
        ref = reflib.remove_connected_components(verts, tris, keep, mins)
        v, t = ops.remove_connected_components(dev(verts), dev(tris), keep, mins)
        assert np.array_equal(v.cpu().numpy(), ref["vertices"]), (keep, mins)
        assert np.array_equal(t.cpu().numpy(), ref["triangles"]), (keep, mins)


def test_kdtree_matches_float32_knn_oracle():
    """Row f-2: k-radius, outlier flags and radius-neighbour counts against the float32
    restatement of the nanoflann queries (oracle/ops_cpu.py)."""
    import adaptivesurfacereconstruction as asr
    from asr_b200 import clouds
    from oracle import ops_cpu
    for name, c in (("blob", clouds.adaptive_blob(20000, seed=5)), ("sphere", clouds.sphere(6000, seed=1))):
        pts = c["points"]
        tree = asr.KDTree(pts)
        for k in (24, 1, 32, 7, 33, 48, 64):  # > 32: the two-list form of the kernel
            r = tree.compute_k_radius(k)
            assert r.dtype == np.float32 and np.array_equal(r, ops_cpu.k_radius(pts, k)), (name, k)
        r24 = tree.compute_k_radius(24)
        rad = (r24 * np.random.default_rng(0).uniform(0.3, 2.0, len(r24))).astype(np.float32)
        got = tree.compute_inlier(rad, 0.5, 24, 1)
        ref = ops_cpu.knn_inlier(pts, rad, 0.5, 24, 1)
        assert got.dtype == np.bool_ and (got != ref).mean() < 1e-3  # ties at the k-th distance pick either point
        assert 0.0 < ref.mean() < 1.0
        got3 = tree.compute_inlier(rad, 0.7, 16, 3)
        assert (got3 != ops_cpu.knn_inlier(pts, rad, 0.7, 16, 3)).mean() < 1e-3
        got4 = tree.compute_inlier(rad, 0.6, 40, 5)
        assert (got4 != ops_cpu.knn_inlier(pts, rad, 0.6, 40, 5)).mean() < 1e-3
        cnt = tree.compute_radius_neighbors(r24)
        assert isinstance(cnt, list) and np.array_equal(np.asarray(cnt, np.int32), ops_cpu.radius_neighbor_counts(pts, r24))
    with pytest.raises(ValueError):
        asr.KDTree(np.zeros((5, 2), np.float32))
    with pytest.raises(ValueError):  # the backend's limit (the reference accepts any k)
        tree.compute_k_radius(65)


def test_reconstruct_surface_estimates_radii():
    import adaptivesurfacereconstruction as asr
    from asr_b200 import clouds, model
    c = clouds.sphere(6000, seed=2)
    net = model.seeded_weights(model.UNet(5), seed=4).cuda()
    mesh = asr.reconstruct_surface(c["points"], c["normals"], None, model=net, contouring_value_threshold=1e9)
    assert set(mesh) == {"vertices", "triangles"} and mesh["triangles"].shape[1] == 3
