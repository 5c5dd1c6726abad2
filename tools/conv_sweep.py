"""Dev tool: time the U-Net conv stack on the bench cloud under different kernel options."""
import sys, time, json
sys.path.insert(0, '/root/repo/adaptive-surface-reconstruction_b200')
import torch
from asr_b200 import _lib, clouds, model, ops, pipeline
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
c = clouds.thingi_like(n, seed=2)
net = model.seeded_weights(model.UNet(6), seed=0).cuda()
dev = {k: torch.from_numpy(c[k]).cuda() for k in ("points", "normals", "radii")}
d, duals, tree = pipeline.build_input_dict(dev["points"], dev["normals"], dev["radii"], c["bb_min"], c["bb_max"], 6)
feats = net.aggregate(d)
def run(label, **opts):
    for k, v in opts.items():
        _lib.set_option(k, v)
    d.pop("_asr_plans", None)  # plans depend on the options (256-pair tiles, output-stationary index)
    for _ in range(2):
        net.unet(feats, d)
    torch.cuda.synchronize()
    _lib.profile_reset(); _lib.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2):
        net.unet(feats, d)
    e1.record(); torch.cuda.synchronize()
    _lib.profile_enable(False)
    prof = _lib.profile_read()
    conv = sum(v["ms"] for k, v in prof.items() if k.startswith("sparse_conv_tile")) / 2
    top = sorted(((v["ms"] / 2, k) for k, v in prof.items() if k.startswith("sparse_conv_tile")), reverse=True)[:12]
    print("%-28s unet %.1f ms  conv tiles %.1f ms  top: %s" % (label, e0.elapsed_time(e1) / 2, conv,
          ", ".join("%s=%.1f" % (k.split("/")[1], m) for m, k in top)), flush=True)
run("stages=2 mt=1", tc_stages=2, tc_row_groups=1)
if len(sys.argv) > 2 and sys.argv[2] == "one": sys.exit(0)
if len(sys.argv) > 2 and sys.argv[2] == "rbs":
    for v in (14, 16, 17, 18):
        run("row block 2^%d" % v, conv_row_block_shift=v)
    sys.exit(0)
if len(sys.argv) > 2 and sys.argv[2] == "os":
    run("hybrid os", sparse_conv_output_stationary=1)
    sys.exit(0)
if len(sys.argv) > 2 and sys.argv[2] == "ntile":
    run("ntile=64 stages=2", tc_ntile=64)
    run("ntile=64 stages=3", tc_ntile=64, tc_stages=3)
    run("ntile=128 stages=3", tc_ntile=128, tc_stages=3)
    sys.exit(0)
for st in (2, 3):
    run("stages=%d mt=2" % st, tc_stages=st, tc_row_groups=2)
for st in (2, 3, 4):
    run("os stages=%d mt=1" % st, tc_stages=st, tc_row_groups=1, sparse_conv_output_stationary=1)
if len(sys.argv) > 2: sys.exit(0)
_lib.set_option("sparse_conv_output_stationary", 0)
ops.SPARSE_CONV_BACKEND = "fp32"
for m in net.modules():
    if hasattr(m, "_packed"): m._packed = {}
run("fp32 FMA kernel")
