"""GPU parity of the float half (search, continuous conv, sparse conv, decode,
list ops) against oracle/ops_cpu.py, through the C ABI.  Index outputs are
bit-exact; float outputs within 1e-4 abs (the north_star tolerance) — in
practice ~1e-6."""
import numpy as np
import pytest
import torch

from helpers import dev

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _scene(n=30000, seed=1, levels=3):
    from asr_b200 import clouds, ops
    c = clouds.adaptive_blob(n, seed=seed)
    t = ops.Octree(dev(c["points"]), dev(c["radii"]), c["bb_min"], c["bb_max"])
    return c, t, t.grids(levels, True)


def test_multi_radius_search_matches_oracle():
    from asr_b200 import ops
    from oracle import ops_cpu
    c, t, grids = _scene()
    q, r = grids[0]["voxel_centers"], grids[0]["voxel_sizes"]
    idx, d2, rs = ops.multi_radius_search(dev(c["points"]), q, r)
    oi, od, ors = ops_cpu.multi_radius_search(c["points"], q.cpu().numpy(), r.cpu().numpy())
    assert np.array_equal(rs.cpu().numpy(), ors)
    assert np.array_equal(idx.cpu().numpy(), oi)
    assert np.array_equal(d2.cpu().numpy(), od)  # same fp32 op order -> identical bits
    assert len(oi) > len(ors)
    sc = ops.scale_compatibility(r, dev(c["radii"]), idx, rs)
    osc = ops_cpu.scale_compatibility(r.cpu().numpy(), c["radii"], oi, ors)
    assert np.abs(sc.cpu().numpy() - osc).max() <= 1e-6


def test_multi_radius_search_edge_cases():
    from asr_b200 import ops
    from oracle import ops_cpu
    rng = np.random.default_rng(0)
    pts = rng.uniform(-1, 1, (5000, 3)).astype(np.float32)
    pts[:50] = pts[50:100]  # exact duplicates -> distance ties
    q = rng.uniform(-1.5, 1.5, (300, 3)).astype(np.float32)  # some queries outside the cloud
    r = rng.uniform(0.0, 0.6, 300).astype(np.float32)
    r[:5] = 0.0
    r[5] = 10.0  # everything
    idx, d2, rs = ops.multi_radius_search(dev(pts), dev(q), dev(r))
    oi, od, ors = ops_cpu.multi_radius_search(pts, q, r)
    assert np.array_equal(rs.cpu().numpy(), ors)
    assert np.array_equal(idx.cpu().numpy(), oi)
    assert np.array_equal(d2.cpu().numpy(), od)
    # empty inputs
    idx, d2, rs = ops.multi_radius_search(dev(pts), torch.zeros((0, 3), device="cuda"), torch.zeros(0, device="cuda"))
    assert idx.numel() == 0 and rs.tolist() == [0]
    idx, d2, rs = ops.multi_radius_search(torch.zeros((0, 3), device="cuda"), dev(q), dev(r))
    assert idx.numel() == 0 and rs.tolist() == [0] * 301


def test_continuous_conv_matches_oracle():
    from asr_b200 import ops
    from oracle import ops_cpu
    c, t, grids = _scene()
    g = grids[0]
    pts = dev(c["points"])
    idx, d2, rs = ops.multi_radius_search(pts, g["voxel_centers"], g["voxel_sizes"])
    sc = ops.scale_compatibility(g["voxel_sizes"], dev(c["radii"]), idx, rs)
    imp = ops.aggregation_importance(sc, d2)
    oimp = torch.from_numpy(ops_cpu.scale_compatibility(g["voxel_sizes"].cpu().numpy(), c["radii"],
                                                        idx.cpu().numpy(), rs.cpu().numpy())) * \
        ops_cpu.window_poly6(d2.cpu())
    assert (imp.cpu() - oimp).abs().max() <= 1e-6
    gen = torch.Generator().manual_seed(0)
    W = (torch.rand((4, 4, 4, 4, 32), generator=gen) - 0.5)
    bias = torch.rand(32, generator=gen) - 0.5
    feats = torch.from_numpy(np.concatenate([c["normals"], np.ones((len(c["normals"]), 1), np.float32)], 1))
    empty = torch.empty(0)
    for use_imp in (True, False):
        ni = imp if use_imp else None
        out = ops.continuous_conv(W.cuda(), g["voxel_centers"], g["voxel_sizes"], torch.zeros(3).cuda(), pts,
                                  feats.cuda(), None, idx, ni, rs, normalize=True)
        ref = ops_cpu.continuous_conv(W, g["voxel_centers"].cpu(), g["voxel_sizes"].cpu(), torch.zeros(3),
                                      torch.from_numpy(c["points"]), feats, empty, idx.cpu(),
                                      imp.cpu() if use_imp else empty, rs.cpu(), normalize=True,
                                      dtype=torch.float64)
        assert (out.cpu().double() - ref).abs().max() <= TOL
        assert ref.abs().max() > 0.1
    fused = ops.continuous_conv(W.cuda(), g["voxel_centers"], g["voxel_sizes"], None, pts, feats.cuda(), None, idx,
                                imp, rs, normalize=True, bias=bias.cuda(), relu=True)
    assert (fused.cpu().double() - torch.relu(ref_with(imp, W, g, c, feats, idx, rs) + bias.double())).abs().max() <= TOL


def ref_with(imp, W, g, c, feats, idx, rs):
    from oracle import ops_cpu
    return ops_cpu.continuous_conv(W, g["voxel_centers"].cpu(), g["voxel_sizes"].cpu(), torch.zeros(3),
                                   torch.from_numpy(c["points"]), feats, torch.empty(0), idx.cpu(), imp.cpu(),
                                   rs.cpu(), normalize=True, dtype=torch.float64)


@pytest.mark.parametrize("cin,cout", [(32, 64), (64, 128), (128, 32), (36, 56), (8, 8), (256, 256)])
@pytest.mark.parametrize("level", [0, 1])
@pytest.mark.parametrize("backend", ["tensor", "fp32"])
def test_sparse_conv_within_grid(cin, cout, level, backend, monkeypatch):
    """The generic entry point asr_sparse_conv (what the open3d:: shim runs): tensor = per-tile pair-major tcgen05
    kernel (3xTF32); fp32 = FMA tile kernel.  The model's own path (gx) is covered by tests/test_gpu_gx.py."""
    from asr_b200 import _lib, ops
    from oracle import ops_cpu
    monkeypatch.setattr(ops, "SPARSE_CONV_BACKEND", "tensor" if backend.startswith("tensor") else backend)
    c, t, grids = _scene(n=12000 if cin * cout > 20000 else 30000)
    g = grids[level]
    V = g["neighbors_row_splits"].shape[0] - 1
    gen = torch.Generator().manual_seed(cin * 1000 + cout)
    W = (torch.rand((55, cin, cout), generator=gen) - 0.5) * 0.2
    x = torch.rand((V, cin), generator=gen) - 0.3
    imp = torch.rand(V, generator=gen)
    imp[::7] = 0.0
    bias = torch.rand(cout, generator=gen) - 0.5
    plan = ops.ConvPlan(g["neighbors_index"], g["neighbors_kernel_index"], g["neighbors_row_splits"], 55)
    cpu = {k: v.cpu() for k, v in g.items()}
    empty = torch.empty(0)
    # plain
    out = ops.sparse_conv(plan, W.cuda(), x.cuda())
    ref = ops_cpu.sparse_conv(W, x, empty, cpu["neighbors_index"], cpu["neighbors_kernel_index"], empty,
                              cpu["neighbors_row_splits"], False, dtype=torch.float64)
    assert (out.cpu().double() - ref).abs().max() <= TOL
    assert ref.abs().max() > 0.05
    # normalize by neighbour count, bias + relu fused
    out = ops.sparse_conv(plan, W.cuda(), x.cuda(), normalize=True, bias=bias.cuda(), relu=True)
    ref = ops_cpu.sparse_conv(W, x, empty, cpu["neighbors_index"], cpu["neighbors_kernel_index"], empty,
                              cpu["neighbors_row_splits"], True, dtype=torch.float64)
    assert (out.cpu().double() - torch.relu(ref + bias.double())).abs().max() <= TOL
    # importance-weighted + normalised (SpecialSparseConv with inp_importance, common_torch.py:124-142)
    nimp = imp[cpu["neighbors_index"].long()]
    out_imp = ops.reduce_subarrays_sum(imp.cuda(), g["neighbors_row_splits"], index=g["neighbors_index"])
    ref_imp = ops_cpu.reduce_subarrays_sum(nimp.double(), cpu["neighbors_row_splits"])
    assert (out_imp.cpu().double() - ref_imp).abs().max() <= 1e-5
    out = ops.sparse_conv(plan, W.cuda(), x.cuda(), inp_importance=imp.cuda(), importance_col=0, normalize=True,
                          normalize_col=0, normalizer=out_imp)
    ref = ops_cpu.sparse_conv(W, x, empty, cpu["neighbors_index"], cpu["neighbors_kernel_index"], nimp,
                              cpu["neighbors_row_splits"], True, dtype=torch.float64)
    assert (out.cpu().double() - ref).abs().max() <= TOL
    # same through the per-entry importance argument of the generic op
    out2 = ops.sparse_conv(plan, W.cuda(), x.cuda(), neighbors_importance=nimp.cuda(), importance_col=0,
                           normalize=True, normalize_col=0, normalizer=out_imp)
    assert (out2.cpu().double() - ref).abs().max() <= TOL
    # fused split conv: leading channels plain, trailing 8 importance-normalised
    if cout >= 16:
        col = cout - 8
        out = ops.sparse_conv(plan, W.cuda(), x.cuda(), inp_importance=imp.cuda(), importance_col=col,
                              normalize=True, normalize_col=col, normalizer=out_imp)
        ra = ops_cpu.sparse_conv(W[:, :, :col], x, empty, cpu["neighbors_index"], cpu["neighbors_kernel_index"],
                                 empty, cpu["neighbors_row_splits"], False, dtype=torch.float64)
        rb = ops_cpu.sparse_conv(W[:, :, col:], x, empty, cpu["neighbors_index"], cpu["neighbors_kernel_index"],
                                 nimp, cpu["neighbors_row_splits"], True, dtype=torch.float64)
        assert (out.cpu().double() - torch.cat([ra, rb], 1)).abs().max() <= TOL


def test_transition_convs_and_invert_neighbors_list():
    from asr_b200 import ops
    from oracle import ops_cpu
    c, t, grids = _scene()
    g0, g1 = grids[0], grids[1]
    V0 = g0["up_neighbors_index"].shape[0]
    V1 = g1["neighbors_row_splits"].shape[0] - 1
    inv = ops.invert_neighbors_list(V1, g0["up_neighbors_index"], g0["up_neighbors_row_splits"],
                                    g0["up_neighbors_kernel_index"])
    ref = ops_cpu.invert_neighbors_list(V1, g0["up_neighbors_index"].cpu(), g0["up_neighbors_row_splits"].cpu(),
                                        g0["up_neighbors_kernel_index"].cpu())
    for a, b in zip(inv, ref):
        assert a.dtype == b.dtype and torch.equal(a.cpu(), b)
    # each coarse voxel receives exactly slots 0..7 or a single slot 8
    rs = ref.neighbors_row_splits
    lens = (rs[1:] - rs[:-1]).numpy()
    assert set(lens.tolist()) <= {1, 8}
    gen = torch.Generator().manual_seed(3)
    empty = torch.empty(0)
    # down conv (rows = coarse voxels)
    W = (torch.rand((9, 64, 128), generator=gen) - 0.5) * 0.2
    x = torch.rand((V0, 64), generator=gen) - 0.3
    plan = ops.ConvPlan(inv.neighbors_index, inv.neighbors_attributes, inv.neighbors_row_splits, 9)
    out = ops.sparse_conv(plan, W.cuda(), x.cuda())
    r = ops_cpu.sparse_conv(W, x, empty, ref.neighbors_index, ref.neighbors_attributes, empty, rs, False,
                            dtype=torch.float64)
    assert (out.cpu().double() - r).abs().max() <= TOL
    # up conv (rows = fine voxels, one entry each)
    W = (torch.rand((9, 128, 64), generator=gen) - 0.5) * 0.2
    x = torch.rand((V1, 128), generator=gen) - 0.3
    plan = ops.ConvPlan(g0["up_neighbors_index"], g0["up_neighbors_kernel_index"], g0["up_neighbors_row_splits"], 9)
    out = ops.sparse_conv(plan, W.cuda(), x.cuda())
    r = ops_cpu.sparse_conv(W, x, empty, g0["up_neighbors_index"].cpu(), g0["up_neighbors_kernel_index"].cpu(),
                            empty, g0["up_neighbors_row_splits"].cpu(), False, dtype=torch.float64)
    assert (out.cpu().double() - r).abs().max() <= TOL
    # generic inversion of a ragged list with float attributes and empty rows
    idx = torch.tensor([3, 0, 3, 2, 0, 0], dtype=torch.int32)
    rs_in = torch.tensor([0, 2, 2, 5, 6], dtype=torch.int64)
    attr = torch.arange(6, dtype=torch.float32)
    a = ops.invert_neighbors_list(5, idx.cuda(), rs_in.cuda(), attr.cuda())
    b = ops_cpu.invert_neighbors_list(5, idx, rs_in, attr)
    for u, v in zip(a, b):
        assert torch.equal(u.cpu(), v)


def test_decode_matches_oracle():
    from asr_b200 import ops
    from oracle import model_cpu
    P = model_cpu.init_params(5, seed=0, stress=True)
    gen = torch.Generator().manual_seed(1)
    code = torch.randn((5003, 32), generator=gen)
    shifts = torch.rand((5003, 3), generator=gen) - 0.5
    ws = [P["dense_decoder1.weight"], P["dense_decoder1.bias"], P["dense_decoder2.weight"],
          P["dense_decoder2.bias"], P["dense_decoder3.weight"]]
    wd = [w.cuda() for w in ws]
    ref = model_cpu.decode(P, shifts.double(), code.double())
    out = ops.decode(shifts.cuda(), code.cuda(), *wd)
    assert (out.cpu().double() - ref).abs().max() <= TOL
    out0 = ops.decode(None, code.cuda(), *wd)
    assert (out0.cpu().double() - model_cpu.decode(P, torch.zeros(5003, 3).double(), code.double())).abs().max() <= TOL
    rv, rg = model_cpu.decode_with_gradient(P, shifts.double(), code.double())
    v, g = ops.decode(shifts.cuda(), code.cuda(), *wd, with_gradient=True)
    assert (v.cpu().double() - rv).abs().max() <= TOL and (g.cpu().double() - rg).abs().max() <= TOL
    scale = torch.rand(5003, generator=gen)
    vs = ops.decode(shifts.cuda(), code.cuda(), *wd, signed_scale=scale.cuda())
    assert (vs.cpu()[:, 0].double() - ref[:, 0] * scale.double()).abs().max() <= TOL
    assert torch.equal(vs[:, 1], out[:, 1])
