#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python tools/debug_gx_unet.py 5 adaptive_blob 20000 ) > gpurun_out/s3_gx_debug.log 2>&1
( timeout 600 python -m pytest tests/test_gpu_geometry.py tests/test_gpu_golden.py -x -q 2>&1 | tail -15 ) > gpurun_out/s3_geom.log 2>&1
( timeout 900 python bench.py --steps 3 --warmup 3 --verify --no-cpu-baseline --backend tensor > gpurun_out/s3_bench_verify.json ) 2> gpurun_out/s3_bench_verify.err
echo done
