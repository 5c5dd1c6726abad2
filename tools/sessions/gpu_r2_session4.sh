#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_gx.py -q 2>&1 | tail -30 ) > gpurun_out/s4_gx.log 2>&1
( timeout 900 python tools/debug_gx_unet.py 5 adaptive_blob 20000 ) > gpurun_out/s4_gx_debug.log 2>&1
( timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s4_bench_gx.json ) 2> gpurun_out/s4_bench_gx.err
( timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --backend tensor > gpurun_out/s4_bench_tensor.json ) 2> gpurun_out/s4_bench_tensor.err
echo done
