#!/bin/bash
mkdir -p gpurun_out
( timeout 100 python tools/gx_odd_shapes.py ) > gpurun_out/s45_odd_shapes.log 2>&1
echo done
