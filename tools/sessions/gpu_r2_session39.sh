#!/bin/bash
# 1 GPU: the config-5 cloud (50 M points, 7 levels) on one GPU with the final build
mkdir -p gpurun_out
( time timeout 1200 python bench.py --workload multi_scan --points 50000000 --levels 7 --seed 3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s39_bench_50m.json ) 2> gpurun_out/s39_bench_50m.err
echo done
