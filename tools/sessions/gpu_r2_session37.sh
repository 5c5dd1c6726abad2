#!/bin/bash
# 1 GPU, end of round: launch list of a steady-state pass, full GPU suite, smoke, bench
mkdir -p gpurun_out
( timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 900 --csv --log-file gpurun_out/s37_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > gpurun_out/s37_launches.out 2>&1
( time timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -25 ) > gpurun_out/s37_pytest.log 2>&1
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/s37_smoke.log 2>&1
( timeout 600 python bench.py > gpurun_out/s37_bench.json ) 2> gpurun_out/s37_bench.err
( timeout 900 python bench.py --impl reference > gpurun_out/s37_bench_ref.json ) 2> gpurun_out/s37_bench_ref.err
echo done
