#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_gx.py -q -x 2>&1 | tail -15 ) > gpurun_out/s7_gx.log 2>&1
( timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --radii analytic > gpurun_out/s7_bench_gx_analytic.json ) 2> gpurun_out/s7_bench_gx_analytic.err
( timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s7_bench_gx.json ) 2> gpurun_out/s7_bench_gx.err
echo done
