#!/bin/bash
# 4 GPUs: sharded path check at world 4, bench N=4 and N=2
mkdir -p gpurun_out
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29650 tests/multi_gpu_worker.py thingi_like 200000 6 ) > gpurun_out/s12_worker.log 2>&1
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/s12_bench_n4.json ) 2> gpurun_out/s12_bench_n4.err
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29652 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s12_bench_n2.json ) 2> gpurun_out/s12_bench_n2.err
( ASR_DEBUG_TIMING=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline ) > gpurun_out/s12_bench_timing.json 2> gpurun_out/s12_bench_timing.err
echo done
