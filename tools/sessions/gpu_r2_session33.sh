#!/bin/bash
# 1 GPU, final evidence: sanitizer on the new conv kernel, ncu launch list + full capture, GPU suite, bench
mkdir -p gpurun_out
( time timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/s33_memcheck_smoke.log 2>&1
( time timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_gx.py -q -x -k "within_grid_conv_matches_oracle and 64-64 or transition or split_first" ) > gpurun_out/s33_memcheck_gx.log 2>&1
( time timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/s33_racecheck_smoke.log 2>&1
( timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 470 -c 520 --csv --log-file gpurun_out/s33_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline ) > gpurun_out/s33_launches.out 2>&1
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:gx_conv_kernel -s 4 -c 16 -o gpurun_out/s33_gx_prof -f python bench.py --steps 1 --warmup 0 --profile-run --no-cpu-baseline ) > gpurun_out/s33_ncu.out 2>&1
ncu -i gpurun_out/s33_gx_prof.ncu-rep --page raw --csv > gpurun_out/s33_gx_prof_raw.csv 2>/dev/null
rm -f gpurun_out/s33_gx_prof.ncu-rep
( time timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -25 ) > gpurun_out/s33_pytest.log 2>&1
( timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/s33_bench.json ) 2> gpurun_out/s33_bench.err
echo done
