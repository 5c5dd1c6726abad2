"""GPU (>= 2 devices): the sharded gx path (asr_b200/shard_gx.py: Z-curve ownership, per-rank plans, peer-store halo
pushes through a symmetric-memory arena) gives the single-GPU result on real hardware — VERDICT r1 item 1c.
Skipped on boxes with one GPU; bench.py --gpus N reports the same comparison as `parity_vs_1gpu`."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("cloud,n,levels", [("adaptive_blob", 60000, 5), ("thingi_like", 200000, 6)])
def test_sharded_path_equals_single_gpu(cloud, n, levels):
    world = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", "29641", os.path.join(HERE, "multi_gpu_worker.py"), cloud, str(n), str(levels)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    line = [l for l in r.stdout.splitlines() if l.startswith("MULTI_GPU_RESULT ")]
    assert line, (r.stdout[-2000:], r.stderr[-4000:])
    res = json.loads(line[0][len("MULTI_GPU_RESULT "):])
    assert res["world"] == world
    for p in res["passes"]:
        assert p["values_max_abs"] <= 1e-6, p   # same kernels, same per-row arithmetic
        assert p["vertex_dual_equal"] and p["vertices_equal"], p
        assert p["vertices"] > 0
