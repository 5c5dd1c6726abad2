#!/bin/bash
# 2 GPUs: multi-GPU tests after the quirk-0 helper refactor
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/s43_pytest_multi.log 2>&1
echo done
