#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python tools/gx_sweep.py 10000000 ) > gpurun_out/s26_sweep.log 2>&1
( timeout 600 python tools/search_stats.py ) > gpurun_out/s26_search.log 2>&1
( time timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -25 ) > gpurun_out/s26_pytest.log 2>&1
( timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/s26_bench.json ) 2> gpurun_out/s26_bench.err
echo done
