#!/bin/bash
# 1 GPU: geometry verification at 10 M points against the compiled reference TUs on the final build
mkdir -p gpurun_out
( time timeout 1500 python bench.py --steps 2 --warmup 3 --verify --no-cpu-baseline > gpurun_out/s38_bench_verify.json ) 2> gpurun_out/s38_bench_verify.err
echo done
