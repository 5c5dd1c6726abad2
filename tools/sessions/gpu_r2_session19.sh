#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_gx.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/s19_pytest.log 2>&1
( timeout 600 python tools/gx_sweep.py 10000000 ) > gpurun_out/s19_sweep.log 2>&1
( timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/s19_bench.json ) 2> gpurun_out/s19_bench.err
echo done
