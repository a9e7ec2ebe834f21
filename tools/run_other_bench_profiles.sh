#!/bin/bash
# launch lists (ncu metrics per kernel) of the diamond and 4-clique bench commands -> gpurun_out/, then
#   python tools/ncu_traffic.py <workload name> <csv>     (here, no GPU needed)
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__cycles_elapsed.avg,sm__inst_executed.avg.per_cycle_elapsed,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
K='tc_hash_kernel|tc_hybrid_kernel|tc_support_kernel|k_diamond_sum|kclique_bitmap_kernel|kclique_warp_edge'
ncu --metrics $M --clock-control none -k regex:"$K" --csv --log-file gpurun_out/r02x_launches_bench_diamond_lj.csv python bench.py --workload diamond --steps 3 --warmup 3 --no-cpu --no-stream > gpurun_out/r02x_ncu_diamond.log 2>&1
ncu --metrics $M --clock-control none -k regex:"$K" --csv --log-file gpurun_out/r02x_launches_bench_clique4_s23.csv python bench.py --workload clique4 --steps 3 --warmup 3 --no-cpu --no-stream > gpurun_out/r02x_ncu_clique4.log 2>&1
