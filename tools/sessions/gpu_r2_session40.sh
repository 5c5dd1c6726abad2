#!/bin/bash
# 1 GPU: full GPU suite and smoke on the last build of the round (kNN k <= 64, per-device attributes)
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -25 ) > gpurun_out/s40_pytest.log 2>&1
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/s40_smoke.log 2>&1
echo done
