#!/bin/bash
# 2 GPUs: multi-GPU tests and bench N=2 (with parity_vs_1gpu) on the final conv kernel
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/s42_pytest_multi.log 2>&1
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29662 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s42_bench_n2.json ) 2> gpurun_out/s42_bench_n2.err
echo done
