#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29650 tests/multi_gpu_worker.py adaptive_blob 60000 5 ) > gpurun_out/s10_worker.log 2>&1
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s10_bench_n2.json ) 2> gpurun_out/s10_bench_n2.err
( timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s10_bench_n1.json ) 2> gpurun_out/s10_bench_n1.err
echo done
