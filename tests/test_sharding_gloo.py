"""CPU, world_size 2, gloo: the multi-GPU decomposition (shard.ShardedOps) gives
the same result as the single-process path, row for row, and really exchanges."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(rank, world, port, q, kind="ShardedOps"):
    for p in (ROOT, os.path.join(ROOT, "adaptive-surface-reconstruction_b200"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import cpu_ops
        from asr_b200 import clouds, model, pipeline, shard
        c = clouds.adaptive_blob(6000, seed=3)
        pts, nrm, rad = (torch.from_numpy(c[k]) for k in ("points", "normals", "radii"))
        net = model.seeded_weights(model.UNet(3), seed=1)
        net.K = cpu_ops
        ref = pipeline.reconstruct_vertices(net, pts, nrm, rad, c["bb_min"], c["bb_max"], contouring_value_threshold=1e9)
        K = getattr(shard, kind)(cpu_ops, min_rows=64)
        net.K = K
        out = pipeline.reconstruct_vertices(net, pts, nrm, rad, c["bb_min"], c["bb_max"], contouring_value_threshold=1e9)
        err = float((out["values"] - ref["values"]).abs().max())
        same_v = bool(torch.equal(out["vertex_dual"], ref["vertex_dual"]))
        verr = float((out["vertices"] - ref["vertices"]).abs().max()) if same_v and out["vertices"].numel() else 0.0
        # the aggregation arrays are this rank's share of the global lists
        rs_ref = ref["input_dict"]["aggregation_row_splits"]
        rs_loc = out["input_dict"]["aggregation_row_splits"]
        if kind == "ShardedOps":
            a, b, n = shard.row_range(ref["values"].shape[0], rank, world)
            share_ok = bool(torch.equal(rs_loc, rs_ref[a:b + 1] - rs_ref[a]))
        else:  # this rank's rows are a few index ranges: the local list holds exactly their pair counts
            rows = K._agg[0].rows
            share_ok = bool(torch.equal(rs_loc[1:] - rs_loc[:-1], (rs_ref[1:] - rs_ref[:-1])[rows])) and \
                0 < rows.shape[0] < ref["values"].shape[0]
        q.put((rank, err, same_v, verr, K.collectives, share_ok, int(ref["values"].shape[0]),
               float(ref["values"].abs().max())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,kind", [(2, "ShardedOps"), (3, "ShardedOps"), (2, "SpatialShardedOps"),
                                        (3, "SpatialShardedOps")])
def test_sharded_pipeline_matches_single_process(world, kind):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400) + world + (10 if kind != "ShardedOps" else 0)
    procs = [ctx.Process(target=_run, args=(r, world, port, q, kind)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, same_v, verr, ncoll, share_ok, v0, vmax in res:
        assert err <= 1e-5, (rank, err)
        assert same_v and verr <= 1e-5
        assert ncoll >= 10, ncoll  # convs + aggregation + decode really went through the collective
        assert share_ok and v0 > 200 and vmax > 1e-3


def test_row_range_partitions_exactly():
    from asr_b200 import shard
    for V in (0, 1, 7, 64, 1000, 12345):
        for world in (1, 2, 3, 8):
            cover = []
            for r in range(world):
                a, b, n = shard.row_range(V, r, world)
                assert 0 <= a <= b <= V and b - a <= n
                cover += list(range(a, b))
            assert cover == list(range(V))
