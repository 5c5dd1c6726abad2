"""GPU box: leaves of the CUDA octree vs the geometry oracle for growing clouds; prints where they differ."""
import sys, os, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "adaptive-surface-reconstruction_b200")]
import torch
from asr_b200 import clouds, ops
from oracle import pipeline_cpu, geomlib

def lev(k):
    return (np.floor(np.log2(k.astype(np.float64))) // 3).astype(int)

res = []
for n, radii in [(1_000_000, "knn"), (3_000_000, "knn"), (10_000_000, "analytic"), (10_000_000, "knn")]:
    c = clouds.thingi_like(n, seed=2)
    p = torch.from_numpy(c["points"]).cuda()
    if radii == "knn":
        c["radii"] = ops.KDTree(p).compute_k_radius(24).cpu().numpy()
    t = ops.Octree(p, torch.from_numpy(c["radii"]).cuda(), c["bb_min"], c["bb_max"])
    lg = t.leaves().cpu().numpy().view(np.uint64)
    kind, Cls = pipeline_cpu.geometry_backend(True)
    lr = Cls(c["points"], c["radii"], c["bb_min"], c["bb_max"], 1.0, 0, 21).leaves()
    lp = geomlib.PortOctree(c["points"], c["radii"], c["bb_min"], c["bb_max"], 1.0, 0, 21).leaves()
    a, b = np.setdiff1d(lg, lr), np.setdiff1d(lr, lg)
    r = {"n": n, "radii": radii, "gpu_leaves": int(len(lg)), "ref_leaves": int(len(lr)), "port_leaves": int(len(lp)),
         "ref_kind": kind, "port_equals_ref": bool(np.array_equal(lp, lr)), "gpu_equals_ref": bool(np.array_equal(lg, lr)),
         "gpu_equals_port": bool(np.array_equal(lg, lp)),
         "only_gpu": int(len(a)), "only_ref": int(len(b)), "balance_rounds": t.balance_rounds,
         "only_gpu_levels": np.bincount(lev(a)).tolist() if len(a) else [], "only_ref_levels": np.bincount(lev(b)).tolist() if len(b) else [],
         "only_gpu_first": [hex(int(x)) for x in a[:16]], "only_ref_first": [hex(int(x)) for x in b[:16]]}
    print(json.dumps(r), flush=True)
    res.append(r)
    if len(a) and n >= 10_000_000:
        np.savez_compressed(os.path.join(ROOT, "gpurun_out", "octree_diff_%d_%s.npz" % (n, radii)), only_gpu=a, only_ref=b)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "s2_octree_parity.json"), "w"), indent=1)
