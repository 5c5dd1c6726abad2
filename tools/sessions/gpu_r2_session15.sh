#!/bin/bash
# 8 GPUs: the 10 M bench cloud and config 5 (50 M points, 7 levels) sharded, each with parity against the 1-GPU path
mkdir -p gpurun_out
export ASR_SHARD_ARENA_GB=48
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/s15_bench_n8.json ) 2> gpurun_out/s15_bench_n8.err
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29652 bench.py --gpus 8 --steps 3 --warmup 3 --workload multi_scan --points 50000000 --levels 7 --seed 3 > gpurun_out/s15_bench_50m_n8.json ) 2> gpurun_out/s15_bench_50m_n8.err
echo done
