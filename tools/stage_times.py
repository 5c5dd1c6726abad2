"""Dev tool: stage times of a few consecutive steps (host-synchronised per stage)."""
import sys
sys.path.insert(0, '/root/repo/adaptive-surface-reconstruction_b200')
import torch
from asr_b200 import clouds, model, pipeline
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
c = clouds.thingi_like(n, seed=2)
net = model.seeded_weights(model.UNet(6), seed=0).cuda()
dev = {k: torch.from_numpy(c[k]).cuda() for k in ("points", "normals", "radii")}
for i in range(5):
    tm = pipeline.StageTimer(enabled=(i != 2))
    out = pipeline.reconstruct_vertices(net, dev["points"], dev["normals"], dev["radii"], c["bb_min"], c["bb_max"], timer=tm)
    torch.cuda.synchronize()
    print(i, {k: round(v, 2) for k, v in tm.ms.items()}, flush=True)
    del out
