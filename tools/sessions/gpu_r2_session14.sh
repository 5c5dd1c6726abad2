#!/bin/bash
# 1 GPU: config 5 single-GPU pass with geometry verification against the oracle; compute-sanitizer
mkdir -p gpurun_out
( time timeout 1500 python bench.py --workload multi_scan --points 50000000 --levels 7 --seed 3 --steps 2 --warmup 3 --verify --no-cpu-baseline > gpurun_out/s14_bench_50m.json ) 2> gpurun_out/s14_bench_50m.err
( time timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/s14_memcheck_smoke.log 2>&1
( time timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_gx.py -q -x -k "within_grid_conv_matches_oracle and 64-64 or transition or split_first" ) > gpurun_out/s14_memcheck_gx.log 2>&1
( time timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/s14_racecheck_smoke.log 2>&1
echo done
