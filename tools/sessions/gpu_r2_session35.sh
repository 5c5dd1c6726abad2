#!/bin/bash
# 8 GPUs: bench N=8 at 10 M points (with parity_vs_1gpu) on the final kernels
mkdir -p gpurun_out
export ASR_SHARD_ARENA_GB=48
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29658 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/s35_bench_n8.json ) 2> gpurun_out/s35_bench_n8.err
echo done
