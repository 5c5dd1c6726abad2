"""CPU: the geometry oracles against each other, against the reference-generated
golden fixture, and against known answers derived from the reference code
(SURVEY.md §8c)."""
import os

import numpy as np
import pytest

from oracle import geomlib, reflib

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_sphere3k.npz")
GRID_KEYS = [n for n, _, _ in reflib.GRID_FIELDS]


def _clouds():
    from asr_b200_clouds import clouds
    yield clouds.sphere(20000, seed=0), 5
    yield clouds.adaptive_blob(30000, seed=1), 5
    yield clouds.thingi_like(30000, seed=2), 6
    yield clouds.gaussian_blob(20000, seed=3), 3


@pytest.fixture(scope="module", autouse=True)
def _clouds_module():
    # import the pure-numpy generator module without importing the CUDA package
    import importlib.util
    import sys
    p = os.path.join(os.path.dirname(os.path.dirname(__file__)), "adaptive-surface-reconstruction_b200", "asr_b200",
                     "clouds.py")
    spec = importlib.util.spec_from_file_location("clouds", p)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    pkg = type(sys)("asr_b200_clouds")
    pkg.clouds = m
    sys.modules["asr_b200_clouds"] = pkg
    yield


def test_port_reproduces_golden_fixture():
    g = np.load(GOLD)
    t = geomlib.PortOctree(g["points"], g["radii"], g["bb_min"], g["bb_max"])
    assert np.array_equal(t.leaves(), g["leaves"])
    grids = t.grids(5, True)
    for l, gr in enumerate(grids):
        for k in GRID_KEYS:
            name = "grid%d_%s" % (l, k)
            assert (k in gr) == (name in g.files), name
            if k in gr:
                assert np.array_equal(gr[k], g[name]), name
    assert np.array_equal(t.dual_vertex_indices(), g["dual_vertex_indices"])
    v, vd = geomlib.contour_vertices(g["values"], g["dual_vertex_indices"], g["grid0_voxel_centers"], 1.0)
    # vertex i <-> i-th intersecting dual; the reference appends fan vertices after them
    assert len(v) > 100 and np.array_equal(v, g["mesh_vertices"][:len(v)])


@pytest.mark.skipif(not reflib.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_port_equals_compiled_reference():
    for c, levels in _clouds():
        a = reflib.RefOctree(c["points"], c["radii"], c["bb_min"], c["bb_max"])
        b = geomlib.PortOctree(c["points"], c["radii"], c["bb_min"], c["bb_max"])
        assert np.array_equal(a.nodes(), b.nodes())
        assert np.array_equal(a.leaves(), b.leaves())
        for x, y in zip(a.params(), b.params()):
            assert np.array_equal(x, y)
        for all_info in (True, False):
            for l, (x, y) in enumerate(zip(a.grids(levels, all_info), b.grids(levels, all_info))):
                assert set(x) == set(y), (l, set(x) ^ set(y))
                for k in x:
                    assert np.array_equal(x[k], y[k]), (l, k)
        da = a.dual_vertex_indices()
        assert np.array_equal(da, b.dual_vertex_indices())
        g0 = a.grids(1, True)[0]
        cen, s = g0["voxel_centers"], g0["voxel_sizes"]
        rad = np.linalg.norm(cen - cen.mean(0), axis=1)
        dist = rad - 0.9 * np.median(rad)
        vals = np.stack([dist, np.abs(dist) / s], 1).astype(np.float32)
        m = reflib.create_triangle_mesh(vals, da, cen, 1.0)
        v, vd = geomlib.contour_vertices(vals, da, cen, 1.0)
        assert len(v) > 20 and np.array_equal(m["vertices"][:len(v)], v)


@pytest.mark.skipif(not reflib.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_margin_free_bbox_and_depth_limit_match_reference():
    from asr_b200_clouds import clouds
    c = clouds.sphere(5000, seed=7)
    mn, mx = c["points"].min(0), c["points"].max(0)
    a = reflib.RefOctree(c["points"], c["radii"], mn, mx)
    b = geomlib.PortOctree(c["points"], c["radii"], mn, mx)
    assert list(a.leaves()[:6]) == [2, 3, 4, 5, 6, 7]
    assert np.array_equal(a.leaves(), b.leaves())
    assert np.array_equal(a.dual_vertex_indices(), b.dual_vertex_indices())
    for md in (4, 6):
        a = reflib.RefOctree(c["points"], c["radii"], c["bb_min"], c["bb_max"], max_depth=md)
        b = geomlib.PortOctree(c["points"], c["radii"], c["bb_min"], c["bb_max"], max_depth=md)
        assert np.array_equal(a.leaves(), b.leaves())


def test_known_answers_and_invariants():
    from asr_b200_clouds import clouds
    c = clouds.adaptive_blob(20000, seed=3)
    t = geomlib.PortOctree(c["points"], c["radii"], c["bb_min"], c["bb_max"])
    leaves = t.leaves()
    nodes = t.nodes()
    assert nodes[0] == 1  # root: ComputeKey({0,0,0,0}) == 1 (octreebase.h:59-65)
    assert np.all(np.diff(leaves.astype(np.int64)) > 0) and leaves[0] >= 8
    # every node has all 7 siblings (octree.cpp:110-150)
    sib = nodes[1:].reshape(-1, 8)
    assert np.all(sib[:, 0] % 8 == 0) and np.all(np.diff(sib.astype(np.int64), axis=1) == 1)
    grids = t.grids(5, True)
    for l, g in enumerate(grids):
        rs, idx, slot = g["neighbors_row_splits"], g["neighbors_index"], g["neighbors_kernel_index"]
        V = len(rs) - 1
        assert np.array_equal(idx[rs[:-1]], np.arange(V)) and np.all(slot[rs[:-1]] == 0)  # grid.cpp:102-106
        lens = np.diff(rs)
        assert lens.min() >= 1 and lens.max() <= 25
        inner = np.ones(len(idx), bool)
        inner[rs[:-1]] = False
        assert np.all(np.diff(slot.astype(np.int32))[inner[1:]] > 0)  # slots strictly increase within a row
        finer = ((slot >= 7) & (slot <= 30)).sum()
        coarser = (slot >= 31).sum()
        assert finer == coarser  # adjacency is symmetric
        if l < 4:
            assert np.array_equal(g["up_neighbors_row_splits"], np.arange(V + 1))  # grid.cpp:206-207
            us = g["up_neighbors_kernel_index"]
            Vn = len(grids[l + 1]["neighbors_row_splits"]) - 1
            assert 8 * ((us < 8).sum() // 8) + (us == 8).sum() == V and (us < 8).sum() // 8 + (us == 8).sum() == Vn
    d = t.dual_vertex_indices()
    assert d.max() < len(leaves)
    uniq = np.array([len(set(r)) for r in d[:2000]])
    assert uniq.min() >= 5
    # Morton3d(3,5,7) == 431 (zindex.h:34): key of cell (3,5,7) at level 3 is 431 | 1<<9
    import ctypes
    assert 431 | (1 << 9) == 943
