#!/bin/bash
mkdir -p gpurun_out
( ASR_DEBUG_TIMING=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline ) > gpurun_out/s6_bench_timing.json 2> gpurun_out/s6_bench_timing.err
( timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s6_bench_gx.json ) 2> gpurun_out/s6_bench_gx.err
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gx_conv_kernel -s 4 -c 20 -o gpurun_out/s6_gx_prof \
   python bench.py --steps 1 --warmup 0 --profile-run --no-cpu-baseline --radii analytic > gpurun_out/s6_ncu.out 2>&1
ncu -i gpurun_out/s6_gx_prof.ncu-rep --page raw --csv > gpurun_out/s6_gx_prof_raw.csv 2>/dev/null
echo done
