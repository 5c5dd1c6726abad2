"""GPU box: per-role wait cycles of the gx convolution kernel (dev option gx_trace) on the bench cloud's tables.
Needs the instrumented library: touch csrc/spconv_gx.cu && make -C adaptive-surface-reconstruction_b200/csrc TRACE=1"""
import sys, os, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "adaptive-surface-reconstruction_b200")]
import torch
from asr_b200 import clouds, ops, gx, _lib

NAMES = ["kernel", "gather:empty", "mma:full", "mma:tmem_empty", "epi0:rare_ready", "epi0:tmem_full", "epi0:columns",
         "epi0:copy_out", "epi0:loop", "rare0:staging_free", "rare0:chunk_ready", "rare0:loop", "ring:slot_free",
         "epi4:tmem_full", "mma:loop", "gather:loop"]
c = clouds.thingi_like(10_000_000, seed=2)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
t = ops.Octree(dev(c["points"]), dev(c["radii"]), c["bb_min"], c["bb_max"])
grids = t.grids(3, True)
gen = torch.Generator().manual_seed(0)
for lev, cin, cout in ((0, 64, 64), (0, 32, 32), (1, 128, 128), (2, 256, 128)):
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
    for mode, label in ((1, "stationary pass"), (2, "pair-major pass")):
        _lib.set_option("gx_trace", mode)
        gx.conv(plan, x, f, out=out, scratch=sc)
        torch.cuda.synchronize()
        buf = (C.c_uint * (16 * 148))()
        _lib.check(_lib.lib().asr_gx_trace(None, 148, buf))
        a = np.frombuffer(buf, dtype=np.uint32).reshape(148, 16).astype(np.float64)
        m = a.mean(0)
        print("L%d %dx%d %s: kernel %.0f kcycles/CTA; share of it: " % (lev, cin, cout, label, m[0] / 1e3) +
              ", ".join("%s %.0f%%" % (NAMES[i], 100 * m[i] / m[0]) for i in range(1, 16)), flush=True)
    _lib.set_option("gx_trace", 0)
