"""GPU: parity at the sizes BASELINE.json names ("next" item 1).

config 1 — 100 k-point sphere, 3 grid levels; config 2 — 1 M-point Gaussian blob with k = 24 kNN
radii, 5 grid levels, full v0 network.  The whole path on the GPU against the CPU oracle pipeline
(geometry = the compiled reference TUs where oracle/_ref exists, else the port; network evaluated in
float64): every index array bit-exact, SDF values within 1e-4 abs, and the max abs error of every
intermediate feature tensor of the U-Net (per level) is measured and written to
gpurun_out/r2_config_parity.json.  Plus: run-to-run spread of the GPU path on the same input.
"""
import json
import os

import numpy as np
import pytest
import torch

from helpers import dev

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-4


def _record(name, payload):
    d = os.path.join(ROOT, "gpurun_out")
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, "r2_config_parity.json")
    cur = {}
    if os.path.exists(path):
        try:
            cur = json.load(open(path))
        except Exception:
            cur = {}
    cur[name] = payload
    with open(path, "w") as f:
        json.dump(cur, f, indent=1, sort_keys=True)


def _config(name):
    from asr_b200 import clouds, ops
    from oracle import ops_cpu
    if name == "config1":
        return clouds.sphere(100_000, seed=0), 3, {}
    c = clouds.gaussian_blob(1_000_000, seed=1)
    # radii = distance to the 24th nearest neighbour (itself included,
    # nsearch.cpp:38-48) — computed by the GPU kNN (row f-2) and checked against the float32 oracle
    r = ops.KDTree(dev(c["points"])).compute_k_radius(24).cpu().numpy()
    r_ref = ops_cpu.k_radius(c["points"], 24)
    info = {"knn_radius_bit_identical": bool(np.array_equal(r, r_ref)),
            "knn_radius_max_abs_diff": float(np.abs(r - r_ref).max())}
    assert info["knn_radius_max_abs_diff"] <= 1e-6 * float(r_ref.max())
    c["radii"] = r
    return c, 5, info


@pytest.mark.parametrize("name", ["config1", "config2"])
def test_baseline_config_parity(name):
    from asr_b200 import model, pipeline
    from oracle import geomlib, model_cpu, pipeline_cpu
    c, levels, info = _config(name)
    P = model_cpu.init_params(levels, seed=0, stress=True)
    net = model.from_state_dict(P, levels)
    out = pipeline.reconstruct_vertices(net, dev(c["points"]), dev(c["normals"]), dev(c["radii"]), c["bb_min"],
                                        c["bb_max"])
    times = {}
    rd, rduals = pipeline_cpu.build_input_dict(c, levels, times=times)
    d = out["input_dict"]
    # ---- integer half: the input dict is bit-identical to the one the reference builds
    for k, v in rd.items():
        assert torch.equal(d[k].cpu(), v), k
    assert np.array_equal(out["dual_vertex_indices"].cpu().numpy().astype(np.uint64), rduals)
    # ---- float half, tensor by tensor (oracle evaluated in float64)
    taps, rtaps = {}, {}
    feats, imp = net.aggregate(d)
    code = net.unet((feats, imp), d, taps=taps)
    with torch.no_grad():
        rfeats, rimp = model_cpu.aggregate(P, rd, dtype=torch.float64)
        rcode = model_cpu.unet(P, (rfeats, rimp), rd, levels, dtype=torch.float64, taps=rtaps)
        rvalues = model_cpu.decode(P, torch.zeros(rcode.shape[0], 3), rcode).clone()
        rvalues[:, 0] *= rd["voxel_sizes0"].double()
        # the float32 CPU evaluation of the same network: how far plain fp32 arithmetic itself is from fp64
        taps32 = {}
        f32 = model_cpu.aggregate(P, rd)
        model_cpu.unet(P, f32, rd, levels, taps=taps32)
    err = {"aggregate": {"max_abs_err": float((feats.cpu().double() - rfeats).abs().max()),
                         "max_abs": float(rfeats.abs().max())}}
    for k in sorted(rtaps):
        err[k] = {"max_abs_err": float((taps[k].cpu().double() - rtaps[k]).abs().max()),
                  "cpu_fp32_max_abs_err": float((taps32[k].double() - rtaps[k]).abs().max()),
                  "max_abs": float(rtaps[k].abs().max()), "rows": int(rtaps[k].shape[0]),
                  "channels": int(rtaps[k].shape[1])}
    values = out["values"].cpu().double()
    err["values"] = {"max_abs_err": float((values - rvalues).abs().max()), "max_abs": float(rvalues.abs().max())}
    # ---- contouring of the SAME values is bit-exact (which duals, order, positions)
    vals = out["values"].cpu().numpy()
    pv, pd = geomlib.contour_vertices(vals, rduals, rd["voxel_centers0"].numpy(), 1.0)
    vertex_exact = bool(np.array_equal(out["vertex_dual"].cpu().numpy().astype(np.uint64), pd) and
                        np.array_equal(out["vertices"].cpu().numpy(), pv))
    # end to end against the oracle's own values: duals whose decision flips sit at |value| < TOL
    rv, rvd = geomlib.contour_vertices(rvalues.float().numpy(), rduals, rd["voxel_centers0"].numpy(), 1.0)
    a, b = set(out["vertex_dual"].cpu().numpy().tolist()), set(rvd.tolist())
    payload = {"points": int(c["points"].shape[0]), "levels": levels, "geometry_oracle": times.get("geometry"),
               "V": [int(rd["neighbors_row_splits%d" % i].shape[0] - 1) for i in range(levels)],
               "pairs": int(rd["aggregation_neighbors_index"].shape[0]), "duals": int(rduals.shape[0]),
               "index_arrays_bit_exact": True, "errors_vs_fp64_oracle": err,
               "vertices_bit_exact_given_values": vertex_exact, "vertices": int(len(a)),
               "vertex_duals_only_on_one_side": int(len(a ^ b)), **info}
    _record(name, payload)
    print(json.dumps(payload))
    assert vertex_exact
    assert err["aggregate"]["max_abs_err"] <= TOL
    assert err["values"]["max_abs_err"] <= TOL
    for k in rtaps:
        # per-voxel features: 1e-4 abs (north star) for O(1) features; the seeded "stress" weights drive some
        # levels to |x| ~ 10-40, where the bound is 1e-4 relative to the tensor's magnitude — measured errors
        # (recorded above) are ~10x below it and of the order of the CPU's own fp32-vs-fp64 error
        assert err[k]["max_abs_err"] <= TOL * max(1.0, err[k]["max_abs"]), (k, err[k])
    assert len(a ^ b) <= 0.01 * max(len(a), 1)


def test_run_to_run_spread():
    """same input twice -> max |dSDF| and the number of vertex duals that differ."""
    from asr_b200 import clouds, model, pipeline
    from oracle import model_cpu
    c = clouds.thingi_like(300_000, seed=4)
    net = model.from_state_dict(model_cpu.init_params(6, seed=0, stress=True), 6)
    args = (dev(c["points"]), dev(c["normals"]), dev(c["radii"]), c["bb_min"], c["bb_max"])
    runs = [pipeline.reconstruct_vertices(net, *args) for _ in range(3)]
    v0 = runs[0]["values"]
    spread = max(float((r["values"] - v0).abs().max()) for r in runs[1:])
    d0 = set(runs[0]["vertex_dual"].cpu().numpy().tolist())
    flips = max(len(d0 ^ set(r["vertex_dual"].cpu().numpy().tolist())) for r in runs[1:])
    for k, v in runs[0]["input_dict"].items():
        if isinstance(v, torch.Tensor) and not v.dtype.is_floating_point:
            assert torch.equal(v, runs[1]["input_dict"][k]), k
    payload = {"points": 300_000, "levels": 6, "max_abs_dSDF_between_runs": spread,
               "vertex_duals_differing_between_runs": flips, "vertices": len(d0)}
    _record("run_to_run", payload)
    print(json.dumps(payload))
    assert spread <= 1e-5
    assert flips <= 0.001 * max(len(d0), 1)
