"""GPU: the split-half ("gx") sparse convolution path (csrc/spconv_gx.cu, asr_b200/gx.py) against the
float64 oracle of the Open3D ops (oracle/ops_cpu.py: sparse_conv as SpecialSparseConv.forward calls it,
models/common_torch.py:95-148), table by table and shape by shape, then the whole U-Net."""
import numpy as np
import pytest
import torch

from helpers import dev

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _tables(levels=3, n=25000, seed=3, cloud="adaptive_blob"):
    from asr_b200 import clouds, ops
    c = clouds.make(cloud, n, seed=seed)
    t = ops.Octree(dev(c["points"]), dev(c["radii"]), c["bb_min"], c["bb_max"])
    return t.grids(levels, True)


def _split_roundtrip_error(x):
    from asr_b200 import gx
    return (gx.from_f32(x).to_f32() - x).abs().max().item()


def test_split_half_format_roundtrip():
    from asr_b200 import gx
    g = torch.Generator().manual_seed(0)
    x = (torch.randn((1000, 64), generator=g) * torch.logspace(-6, 3, 64)[None, :]).cuda()
    y = gx.from_f32(x).to_f32()
    big = x.abs() >= 0.125
    assert ((y - x).abs()[big] <= x.abs()[big] * 2.0 ** -21).all()  # >= 22 significant bits
    assert (y - x).abs()[~big].max() <= 2.0 ** -24                  # absolute below
    s = torch.rand(1000, generator=g).cuda()
    z = gx.scale_rows(gx.from_f32(x), s).to_f32()
    assert (z - y * s[:, None]).abs().max() <= 2.0 ** -20 * (y.abs().max() + 1)
    assert not gx.overflow()
    gx.from_f32(torch.full((8, 32), 1e6, device="cuda"))
    assert gx.overflow() and not gx.overflow()  # saturation is reported once, then cleared


# (64, 48) / (128, 96): output rows whose 16-byte chunk count is not a power of two (general copy-out path)
@pytest.mark.parametrize("cin,cout", [(32, 64), (64, 64), (64, 32), (32, 32), (128, 128), (384, 128), (256, 256),
                                      (64, 48), (128, 96)])
def test_within_grid_conv_matches_oracle(cin, cout):
    from asr_b200 import gx
    from oracle import ops_cpu
    g = _tables()[0]
    V = g["neighbors_row_splits"].shape[0] - 1
    gen = torch.Generator().manual_seed(cin * 1000 + cout)
    W = (torch.rand((55, cin, cout), generator=gen) - 0.5) * (2.0 / np.sqrt(7.7 * cin))
    b = (torch.rand(cout, generator=gen) - 0.5) * 0.2
    x = torch.randn((V, cin), generator=gen)
    plan = gx.Plan(g["neighbors_index"], g["neighbors_kernel_index"], g["neighbors_row_splits"], V, 55, gx.MODE_STATIONARY)
    plan.finish()
    assert 0 < plan.num_rare < g["neighbors_index"].shape[0]
    xs = gx.from_f32(x.cuda())
    cpu = {k: v.cpu() for k, v in g.items()}
    ref = ops_cpu.sparse_conv(W, xs.to_f32().cpu(), torch.empty(0), cpu["neighbors_index"], cpu["neighbors_kernel_index"],
                              torch.empty(0), cpu["neighbors_row_splits"], False, dtype=torch.float64)
    y = gx.conv(plan, xs, gx.filter_bank(W.cuda(), b.cuda()), relu=True).to_f32()
    want = torch.relu(ref + b.double())
    assert (y.cpu().double() - want).abs().max() <= TOL
    # fp32 output, no activation, a residual
    r = torch.randn((V, cout), generator=gen)
    rs = gx.from_f32(r.cuda())
    out = torch.empty((V, cout), dtype=torch.float32, device="cuda")
    gx.conv(plan, xs, gx.filter_bank(W.cuda(), b.cuda()), relu=False, res=rs, out_f32=out)
    assert (out.cpu().double() - (ref + b.double() + rs.to_f32().cpu().double())).abs().max() <= TOL
    # bit-reproducible: no atomics anywhere on this path; as column groups (what the model does for wide banks)
    y2 = gx.conv(plan, xs, gx.filter_bank(W.cuda(), b.cuda()), relu=True).to_f32()
    assert torch.equal(y, y2)
    y3 = gx.conv(plan, xs, gx.filter_bank(W.cuda(), b.cuda(), max_cols=max(16, cout // 2)), relu=True).to_f32()
    assert (y3.cpu().double() - want).abs().max() <= TOL


def test_split_first_conv_with_importance_matches_oracle():
    """conv1a | conv1b of the encoder blocks (net_definitions_torch.py:199-210,280-283): the last 8 channels are
    the importance-weighted, importance-normalised convolution."""
    from asr_b200 import gx, model
    from oracle import ops_cpu
    g = _tables()[0]
    V = g["neighbors_row_splits"].shape[0] - 1
    gen = torch.Generator().manual_seed(5)
    blk = model._Block(64, 64, 55, True, 1)
    with torch.no_grad():
        for p in blk.parameters():
            p.copy_((torch.rand(p.shape, generator=gen) - 0.5) * 0.1)
    blk = blk.cuda()
    x = torch.randn((V, 64), generator=gen)
    imp = torch.rand(V + 77, generator=gen)  # longer than V (quirk 0)
    plan = gx.Plan(g["neighbors_index"], g["neighbors_kernel_index"], g["neighbors_row_splits"], V, 55, gx.MODE_STATIONARY)
    xs = gx.from_f32(x.cuda())
    y, out_imp = gx.run_block(blk, xs, plan, imp.cuda(), gx.Scratch())
    cpu = {k: v.cpu() for k, v in g.items()}
    xr = xs.to_f32().cpu()
    nimp = imp[cpu["neighbors_index"].long()]
    a = ops_cpu.sparse_conv(blk.conv1a.kernel.cpu(), xr, torch.empty(0), cpu["neighbors_index"], cpu["neighbors_kernel_index"],
                            torch.empty(0), cpu["neighbors_row_splits"], False, dtype=torch.float64)
    bb = ops_cpu.sparse_conv(blk.conv1b.kernel.cpu(), xr, torch.empty(0), cpu["neighbors_index"], cpu["neighbors_kernel_index"],
                             nimp, cpu["neighbors_row_splits"], True, dtype=torch.float64)
    want = torch.relu(torch.cat([a + blk.conv1a.bias.cpu().double(), bb + blk.conv1b.bias.cpu().double()], 1))
    assert (y.to_f32().cpu().double() - want).abs().max() <= TOL
    assert (out_imp.cpu() - ops_cpu.reduce_subarrays_sum(nimp, cpu["neighbors_row_splits"])).abs().max() <= 1e-4
    # importances spread over 12 decades (voxels far from any point): the normalised channels are a weighted MEAN,
    # so they must stay accurate however small the weights are — the importance is applied in fp32 in the epilogue
    imp2 = imp * torch.pow(10.0, -12 * torch.rand(imp.shape, generator=gen))
    y2, _ = gx.run_block(blk, xs, plan, imp2.cuda(), gx.Scratch())
    nimp2 = imp2[cpu["neighbors_index"].long()]
    bb2 = ops_cpu.sparse_conv(blk.conv1b.kernel.cpu().double(), xr.double(), torch.empty(0), cpu["neighbors_index"],
                              cpu["neighbors_kernel_index"], nimp2.double(), cpu["neighbors_row_splits"], True,
                              dtype=torch.float64)
    want2 = torch.relu(bb2 + blk.conv1b.bias.cpu().double())
    assert (y2.to_f32().cpu().double()[:, 56:] - want2).abs().max() <= TOL


@pytest.mark.parametrize("cin,cout", [(64, 128), (256, 256)])
def test_transition_convs_match_oracle(cin, cout):
    """K = 9 tables: down (inverted up table, output-stationary with 8 dense child slots) and up (one entry per row,
    pair-major with the final epilogue)."""
    from asr_b200 import gx, ops
    from oracle import ops_cpu
    grids = _tables()
    g0, g1 = grids[0], grids[1]
    V0, V1 = g0["neighbors_row_splits"].shape[0] - 1, g1["neighbors_row_splits"].shape[0] - 1
    gen = torch.Generator().manual_seed(cin + cout)
    W = (torch.rand((9, cin, cout), generator=gen) - 0.5) * (2.0 / np.sqrt(cin))
    b = (torch.rand(cout, generator=gen) - 0.5) * 0.2
    ui, uk, us = g0["up_neighbors_index"], g0["up_neighbors_kernel_index"], g0["up_neighbors_row_splits"]
    # down: rows = coarse voxels, inputs = fine voxels
    inv = ops.invert_neighbors_list(V1, ui, us, uk)
    plan = gx.Plan(inv.neighbors_index, inv.neighbors_attributes, inv.neighbors_row_splits, V0, 9, gx.MODE_STATIONARY)
    x = gx.from_f32(torch.randn((V0, cin), generator=gen).cuda())
    y = gx.conv(plan, x, gx.filter_bank(W.cuda(), b.cuda()), relu=True).to_f32()
    ref = ops_cpu.sparse_conv(W, x.to_f32().cpu(), torch.empty(0), inv.neighbors_index.cpu(), inv.neighbors_attributes.cpu(),
                              torch.empty(0), inv.neighbors_row_splits.cpu(), False, dtype=torch.float64)
    assert y.shape == (V1, cout)
    assert (y.cpu().double() - torch.relu(ref + b.double())).abs().max() <= TOL
    # up: rows = fine voxels, inputs = coarse voxels; written into a channel slice of a wider buffer
    W2 = (torch.rand((9, cout, cin), generator=gen) - 0.5) * (2.0 / np.sqrt(cout))
    plan_up = gx.Plan(ui, uk, us, V1, 9, gx.MODE_PAIR_FINAL)
    xc = gx.from_f32(torch.randn((V1, cout), generator=gen).cuda())
    wide = gx.H2.empty(V0, cin + 64, "cuda")
    wide.buf[:V0].fill_(7.0)
    gx.conv(plan_up, xc, gx.filter_bank(W2.cuda()), relu=True, out=wide.slice(0, cin))
    ref = ops_cpu.sparse_conv(W2, xc.to_f32().cpu(), torch.empty(0), ui.cpu(), uk.cpu(), torch.empty(0), us.cpu(), False,
                              dtype=torch.float64)
    assert (wide.slice(0, cin).to_f32().cpu().double() - torch.relu(ref)).abs().max() <= TOL
    assert (wide.slice(cin, 64).to_f32() == 14.0).all()  # the neighbouring slice is untouched (hi 7 + lo 7)


@pytest.mark.parametrize("levels,cloud,n", [(5, "adaptive_blob", 20000), (3, "sphere", 30000), (6, "thingi_like", 40000)])
def test_unet_on_gx_matches_oracle(levels, cloud, n, monkeypatch):
    from asr_b200 import clouds, model, ops, pipeline
    from oracle import model_cpu, pipeline_cpu
    monkeypatch.setattr(ops, "SPARSE_CONV_BACKEND", "gx")
    c = clouds.make(cloud, n, seed=2)
    P = model_cpu.init_params(levels, seed=0, stress=True)
    net = model.from_state_dict(P, levels)
    out = pipeline.reconstruct_vertices(net, dev(c["points"]), dev(c["normals"]), dev(c["radii"]), c["bb_min"], c["bb_max"])
    d = out["input_dict"]
    rd, _ = pipeline_cpu.build_input_dict(c, levels)
    taps, rtaps = {}, {}
    feats, imp = net.aggregate(d)
    code = net.unet((feats, imp), d, taps=taps)
    with torch.no_grad():
        rf = model_cpu.aggregate(P, rd, dtype=torch.float64)
        rcode = model_cpu.unet(P, rf, rd, levels, dtype=torch.float64, taps=rtaps)
        rvalues = model_cpu.decode(P, torch.zeros(rcode.shape[0], 3), rcode).clone()
        rvalues[:, 0] *= rd["voxel_sizes0"].double()
    for k in sorted(rtaps):
        err = (taps[k].cpu().double() - rtaps[k]).abs().max().item()
        assert err <= TOL * max(1.0, rtaps[k].abs().max().item()), (k, err)
    assert (code.cpu().double() - rcode).abs().max() <= TOL * max(1.0, rcode.abs().max().item())
    assert (out["values"].cpu().double() - rvalues).abs().max() <= TOL
    # same result from the round-1 pair-major kernels (fp32 activations, 3xTF32)
    monkeypatch.setattr(ops, "SPARSE_CONV_BACKEND", "tensor")
    code_tc = net.unet((feats, imp), d)
    assert (code_tc - code).abs().max().item() <= TOL * max(1.0, rcode.abs().max().item())


def test_dev_options_and_trace_guard():
    """unknown options are refused; the per-role trace of the gx kernel exists only in the instrumented build
    (make TRACE=1) and says so instead of returning stale counters"""
    import ctypes
    from asr_b200 import _lib
    with pytest.raises(ValueError):
        _lib.set_option("no_such_option", 1)
    for name in ("gx_acc_groups", "gx_tma_gather", "gx_l1_gather", "gx_max_stages", "gx_ablate", "gx_single_tmem",
                 "gx_one_team", "gx_trace"):
        _lib.set_option(name, 0)
    buf = (ctypes.c_uint * 16)()
    rc = _lib.lib().asr_gx_trace(None, 1, buf)
    if rc != 0:
        assert b"TRACE=1" in _lib.lib().asr_last_error()
