"""Dev tool: one U-Net pass with the output-stationary kernel on (for ncu)."""
import sys
sys.path.insert(0, '/root/repo/adaptive-surface-reconstruction_b200')
import torch
from asr_b200 import _lib, clouds, model, ops, pipeline
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
_lib.set_option("sparse_conv_output_stationary", 1)
c = clouds.thingi_like(n, seed=2)
net = model.seeded_weights(model.UNet(6), seed=0).cuda()
dev = {k: torch.from_numpy(c[k]).cuda() for k in ("points", "normals", "radii")}
d, duals, tree = pipeline.build_input_dict(dev["points"], dev["normals"], dev["radii"], c["bb_min"], c["bb_max"], 6)
feats = net.aggregate(d)
net.unet(feats, d)
torch.cuda.synchronize()
