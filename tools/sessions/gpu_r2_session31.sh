#!/bin/bash
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 ) > gpurun_out/s31_pytest.log 2>&1
( timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/s31_bench.json ) 2> gpurun_out/s31_bench.err
echo done
