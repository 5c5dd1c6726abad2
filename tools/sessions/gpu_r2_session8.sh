#!/bin/bash
mkdir -p gpurun_out
( ASR_DEBUG_TIMING=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > gpurun_out/s8_bench_timing.json 2> gpurun_out/s8_bench_timing.err
( timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s8_bench_gx.json ) 2> gpurun_out/s8_bench_gx.err
( timeout 900 python -m pytest tests/test_gpu_gx.py -q -x 2>&1 | tail -5 ) > gpurun_out/s8_gx.log 2>&1
echo done
