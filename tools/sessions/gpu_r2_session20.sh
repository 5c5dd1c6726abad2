#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gx_conv_kernel -c 4 -o gpurun_out/s20_gx_prof -f python tools/gx_prof.py > gpurun_out/s20_ncu.out 2>&1
ls -la gpurun_out/s20_gx_prof.ncu-rep >> gpurun_out/s20_ncu.out
echo done
