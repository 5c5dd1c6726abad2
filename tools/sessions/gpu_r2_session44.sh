#!/bin/bash
# 1 GPU: smoke on the last build, and the bench on round 1's cloud (analytic radii) for a like-for-like conv-stack number
mkdir -p gpurun_out
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/s44_smoke.log 2>&1
( timeout 400 python bench.py --radii analytic --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s44_bench_analytic.json ) 2> gpurun_out/s44_bench_analytic.err
echo done
