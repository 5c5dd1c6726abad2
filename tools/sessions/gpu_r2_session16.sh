#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python tools/gx_sweep.py 10000000 ) > gpurun_out/s16_sweep.log 2>&1
echo done
