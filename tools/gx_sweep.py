"""GPU box: time single gx convolutions (CUDA events) on the tables of the bench cloud under dev knobs."""
import sys, os, json, itertools
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "adaptive-surface-reconstruction_b200")]
import torch
from asr_b200 import clouds, ops, gx, _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
c = clouds.thingi_like(n, seed=2)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
t = ops.Octree(dev(c["points"]), dev(c["radii"]), c["bb_min"], c["bb_max"])
grids = t.grids(3, True)
gen = torch.Generator().manual_seed(0)
shapes = [(0, 64, 64), (0, 32, 32), (1, 128, 128), (2, 256, 128)]
plans = {}
for lev in (0, 1, 2):
    g = grids[lev]
    V = g["neighbors_row_splits"].shape[0] - 1
    plans[lev] = gx.Plan(g["neighbors_index"], g["neighbors_kernel_index"], g["neighbors_row_splits"], V, 55, gx.MODE_STATIONARY).finish()
    print("level", lev, "V", V, "E", g["neighbors_index"].shape[0], "rare", plans[lev].num_rare, flush=True)

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

# gx_ablate (timing only, garbage results): 1 no filter copies, 2 no pair-buffer reads, 4 no row gathers, 8 no stores, 16 no MMAs
configs = [dict()]
for lev, cin, cout in shapes:
    plan = plans[lev]
    V = plan.num_out
    W = ((torch.rand((55, cin, cout), generator=gen) - 0.5) * 0.2).cuda()
    x = gx.from_f32(torch.randn((V, cin), generator=gen).cuda())
    f = gx.filter_bank(W)
    out = gx.H2.empty(V, cout, "cuda")
    sc = gx.Scratch()
    res = {}
    for cfg in configs:
        for k in ("gx_l1_gather", "gx_max_stages", "gx_acc_groups", "gx_ablate", "gx_single_tmem", "gx_one_team"):
            _lib.set_option(k, cfg.get(k, 0))
        res[json.dumps(cfg)] = round(timeit(lambda: gx.conv(plan, x, f, out=out, scratch=sc)), 3)
    print("L%d %dx%d:" % (lev, cin, cout), res, flush=True)
