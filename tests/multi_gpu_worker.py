"""Worker of tests/test_gpu_multi.py: run under torchrun with one rank per GPU.  Every rank runs the sharded gx path
on the same cloud; rank 0 also runs the single-GPU path and prints one JSON line with the differences."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "adaptive-surface-reconstruction_b200")]

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from asr_b200 import clouds, model, ops, pipeline, shard_gx  # noqa: E402


def main():
    cloud_name, n, levels = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    c = clouds.make(cloud_name, n, seed=2)
    net = model.seeded_weights(model.UNet(levels), seed=0).cuda()
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    args = (dev(c["points"]), dev(c["normals"]), dev(c["radii"]), c["bb_min"], c["bb_max"])
    arena = shard_gx.Arena(int(float(os.environ.get("ASR_SHARD_ARENA_GB", "4")) * (1 << 30)), dist.group.WORLD)
    ctx = shard_gx.ShardContext(arena, min_rows=int(os.environ.get("ASR_SHARD_MIN_ROWS", "2000")))
    res = []
    for it in range(2):  # two passes: the arena is reused
        out = shard_gx.reconstruct_vertices(net, ctx, *args)
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            one = pipeline.reconstruct_vertices(net, *args)
            res.append({"values_max_abs": float((one["values"] - out["values"]).abs().max()),
                        "values_equal": bool(torch.equal(one["values"], out["values"])),
                        "vertex_dual_equal": bool(one["vertex_dual"].shape == out["vertex_dual"].shape and
                                                  torch.equal(one["vertex_dual"], out["vertex_dual"])),
                        "vertices_equal": bool(one["vertices"].shape == out["vertices"].shape and
                                               torch.equal(one["vertices"], out["vertices"])),
                        "vertices": int(one["vertices"].shape[0]), "V0": int(one["values"].shape[0]),
                        "owned_rows_rank0": [int(r.shape[0]) for r in ctx.rows], "exchanges": ctx.exchanges})
        dist.barrier()
    if rank == 0:
        print("MULTI_GPU_RESULT " + json.dumps({"world": world, "passes": res}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
