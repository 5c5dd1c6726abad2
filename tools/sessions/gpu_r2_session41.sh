#!/bin/bash
# 1 GPU: the default bench command once more (stage-timing pass stabilised)
mkdir -p gpurun_out
( timeout 600 python bench.py > gpurun_out/s41_bench.json ) 2> gpurun_out/s41_bench.err
echo done
