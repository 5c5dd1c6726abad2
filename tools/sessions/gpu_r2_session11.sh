#!/bin/bash
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gx_conv_kernel -s 4 -c 16 -o gpurun_out/s11_gx_prof \
   python bench.py --steps 1 --warmup 0 --profile-run --no-cpu-baseline --radii analytic > gpurun_out/s11_ncu.out 2>&1
ncu -i gpurun_out/s11_gx_prof.ncu-rep --page raw --csv > gpurun_out/s11_gx_prof_raw.csv 2>/dev/null
echo done
