#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python tools/gx_sweep.py 10000000 ) > gpurun_out/s25_sweep.log 2>&1
( time timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -25 ) > gpurun_out/s25_pytest.log 2>&1
( timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/s25_bench.json ) 2> gpurun_out/s25_bench.err
echo done
