"""GPU box, under ncu: one L0 64->64 and one L1 128->128 gx convolution on the bench cloud's tables."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "adaptive-surface-reconstruction_b200")]
import torch
from asr_b200 import clouds, ops, gx

c = clouds.thingi_like(10_000_000, seed=2)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
t = ops.Octree(dev(c["points"]), dev(c["radii"]), c["bb_min"], c["bb_max"])
grids = t.grids(2, True)
gen = torch.Generator().manual_seed(0)
for lev, cin, cout in ((0, 64, 64), (1, 128, 128)):
    g = grids[lev]
    V = g["neighbors_row_splits"].shape[0] - 1
    plan = gx.Plan(g["neighbors_index"], g["neighbors_kernel_index"], g["neighbors_row_splits"], V, 55, gx.MODE_STATIONARY).finish()
    W = ((torch.rand((55, cin, cout), generator=gen) - 0.5) * 0.2).cuda()
    x = gx.from_f32(torch.randn((V, cin), generator=gen).cuda())
    f = gx.filter_bank(W)
    out = gx.H2.empty(V, cout, "cuda")
    sc = gx.Scratch()
    gx.conv(plan, x, f, out=out, scratch=sc)
    torch.cuda.synchronize()
print("done")
