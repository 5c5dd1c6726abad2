"""GPU box: distribution of the aggregation search's row lengths on the bench cloud and the time of its phases."""
import sys, os, types
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "adaptive-surface-reconstruction_b200")]
import torch
import bench
from asr_b200 import ops, _lib

args = types.SimpleNamespace(workload="thingi_like", seed=2, radii="knn")
c = bench.make_cloud(args, 10_000_000, on_gpu=True)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
pts, rad = dev(c["points"]), dev(c["radii"])
t = ops.Octree(pts, rad, c["bb_min"], c["bb_max"])
g = t.grids(1, True)[0]
q, r = g["voxel_centers"], g["voxel_sizes"]
for rep in range(2):
    _lib.profile_reset(); _lib.profile_enable(True)
    idx, d2, rs = ops.multi_radius_search(pts, q, r, frame=t.search_frame())
    torch.cuda.synchronize()
    prof = _lib.profile_read(); _lib.profile_enable(False)
print({k: round(v["ms"], 3) for k, v in prof.items()})
n = (rs[1:] - rs[:-1]).cpu().numpy()
print("queries", n.size, "pairs", int(n.sum()), "mean", n.mean(), "max", n.max())
for th in (32, 64, 128, 256, 512, 1024, 2048):
    m = n > th
    print("rows > %d: %d (%.4f %%), pairs in them %d (%.2f %%)" % (th, m.sum(), 100 * m.mean(), n[m].sum(), 100 * n[m].sum() / n.sum()))
print("percentiles 50/90/99/99.9/99.99:", np.percentile(n, [50, 90, 99, 99.9, 99.99]))
