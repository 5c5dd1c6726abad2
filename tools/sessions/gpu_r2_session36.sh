#!/bin/bash
# 4 GPUs: bench N=4 at 10 M points (with parity_vs_1gpu) on the final kernels
mkdir -p gpurun_out
export ASR_SHARD_ARENA_GB=48
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29659 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/s36_bench_n4.json ) 2> gpurun_out/s36_bench_n4.err
echo done
