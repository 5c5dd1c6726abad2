#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python tools/gx_trace.py ) > gpurun_out/s24_trace.log 2>&1
echo done
