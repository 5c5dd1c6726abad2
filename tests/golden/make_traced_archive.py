"""Writes tests/golden/_local/ref_unet5_traced.pt — a TorchScript archive made the way the
reference makes its `model.pt` (models/v0/convert_tf2torchscript.py:110-122): the reference's OWN
`UNet5` (imported from /root/reference) is driven through aggregate / unet / decode and
`torch.jit.trace_module`d, then saved.  Run in the build container, where /root/reference exists:

    python tests/golden/make_traced_archive.py

Open3D is not installed, so the `open3d::*` ops the graph records are supplied by oracle/o3d_shim
registered under the real namespace (ASR_ORACLE_O3D_NAMESPACE=open3d) with Open3D's full schemas —
the archive's graphs therefore carry exactly the calls (`open3d::sparse_conv(..., normalize,
max_temp_mem_MB)`, `open3d::continuous_conv(..., align_corners, coordinate_mapping, normalize,
interpolation, max_temp_mem_MB)`) an archive traced against real Open3D carries.

Inputs and weights are those of tests/golden/ref_sphere3k.npz (seed 11), whose `code` / `values` arrays
are therefore the expected outputs of this archive.  The archive holds the full 92 M-parameter network
(369 MB): it is NOT committed (tests/golden/_local/ is git-ignored) but travels to the GPU box with the
working tree, where tests/test_gpu_shims.py::test_traced_reference_archive_runs_on_the_cuda_shim runs
it on the CUDA ops of this repo.  This process never imports the product package."""
import os
import sys

os.environ["ASR_ORACLE_O3D_NAMESPACE"] = "open3d"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "o3d_shim"))
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402
import torch  # noqa: E402

from models.v0.net_definitions_torch import UNet5  # noqa: E402  (the reference's model)
from oracle import model_cpu  # noqa: E402


def golden_inputs(g):
    inp = {"points": torch.from_numpy(g["points"]),
           "feats": torch.from_numpy(np.concatenate([g["normals"], np.ones((len(g["points"]), 1), np.float32)], 1))}
    for name in g.files:
        if name.startswith("grid") and not name.endswith("voxel_keys"):
            inp[name[6:] + name[4]] = torch.from_numpy(g[name])
    for k in ("aggregation_neighbors_index", "aggregation_neighbors_dist", "aggregation_row_splits",
              "aggregation_scale_compat"):
        inp[k] = torch.from_numpy(g[k])
    return inp


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    g = np.load(os.path.join(here, "ref_sphere3k.npz"))
    data = golden_inputs(g)
    model = UNet5(with_importance="all", normalized_channels=8, residual_skip_connection=True)
    model.load_state_dict(model_cpu.init_params(5, seed=int(g["weights_seed"]), stress=True))
    with torch.no_grad():
        agg = model.aggregate(data)
        code = model.unet(agg, data)
        shift = torch.zeros([code.shape[0], 3])
        script = torch.jit.trace_module(model, {"aggregate": data, "unet": (agg, data), "decode": (shift, code)})
    assert np.abs(code.numpy() - g["code"]).max() <= 1e-6
    os.makedirs(os.path.join(here, "_local"), exist_ok=True)
    path = os.path.join(here, "_local", "ref_unet5_traced.pt")
    script.save(path)
    calls = sorted({n.kind() for m in ("aggregate", "unet", "decode")
                    for n in getattr(script, m).inlined_graph.nodes() if n.kind().startswith("open3d::")})
    print("wrote", path, os.path.getsize(path) >> 20, "MiB; open3d ops in the graphs:", calls)


if __name__ == "__main__":
    main()
