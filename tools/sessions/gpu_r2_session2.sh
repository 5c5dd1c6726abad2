#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_gx.py -x -q -s 2>&1 | tail -60 ) > gpurun_out/s2_gx.log 2>&1
( time timeout 600 python -m pytest tests/test_gpu_shims.py -q -s 2>&1 | tail -30 ) > gpurun_out/s2_shims.log 2>&1
( time timeout 1200 python tools/debug_octree_parity.py ) > gpurun_out/s2_octree.log 2>&1
echo done
