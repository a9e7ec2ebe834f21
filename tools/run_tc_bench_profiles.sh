set -x
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__cycles_elapsed.avg,sm__inst_executed.avg.per_cycle_elapsed,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
for S in 22 24; do
  ncu --metrics $M --clock-control none -k regex:'tc_hash_kernel|tc_hybrid_kernel' --csv --log-file gpurun_out/r02q_launches_bench_tc_s$S.csv python bench.py --scale $S --steps 3 --warmup 3 --no-cpu --no-stream > gpurun_out/r02q_ncu_bench_s$S.log 2>&1
  python bench.py --scale $S --steps 10 --warmup 3 --no-cpu --no-stream > gpurun_out/r02q_tc_s$S.json 2> gpurun_out/r02q_tc_s$S.err
  tail -c 600 gpurun_out/r02q_tc_s$S.json
done
