#!/bin/bash
# round-2 GPU session 1: full GPU test suite (new parity tests included), baseline bench with kNN radii + --verify,
# per-kernel ncu counters of the non-conv stages
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s1_smi.txt 2>&1
nproc >> gpurun_out/s1_smi.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 ) > gpurun_out/s1_pytest.log 2>&1
( time timeout 900 python bench.py --steps 10 --warmup 3 --verify > gpurun_out/s1_bench_knn.json ) 2> gpurun_out/s1_bench_knn.err
( time timeout 600 python bench.py --steps 10 --warmup 3 --radii analytic --no-cpu-baseline > gpurun_out/s1_bench_analytic.json ) 2> gpurun_out/s1_bench_analytic.err
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,launch__registers_per_thread,launch__occupancy_limit_registers
timeout 1200 ncu --metrics $M --clock-control none -k regex:'adjacency|dual_|ball_query|cconv4|decode_thread|contour_|point_group|balance_round|ancestor|leaf_|coarsen|up_table|voxel_info|scale_compat|importance|hash_|cell_|gather_points|point_code|entry_rows|gather_pairs|group_begin|tile_list|conv_epilogue|row_importance' \
   -c 200 --csv --log-file gpurun_out/s1_ncu_nonconv.csv python bench.py --steps 1 --warmup 1 --profile-run --no-cpu-baseline > gpurun_out/s1_ncu_nonconv.out 2>&1
echo done
