set -u
OUT=gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,launch__block_size
timeout -s KILL 420 python bench.py > $OUT/bench_r02_cfg2.json 2> $OUT/bench_cfg2.err; timeout 20 python tools/bench_brief.py $OUT/bench_r02_cfg2.json cfg2
timeout -s KILL 200 python bench.py --workload cfg3 --no-cpu-baseline > $OUT/bench_r02_cfg3.json 2> $OUT/bench_cfg3.err; timeout 20 python tools/bench_brief.py $OUT/bench_r02_cfg3.json cfg3
timeout -s KILL 120 python bench.py --workload cfg1 --no-cpu-baseline > $OUT/bench_r02_cfg1.json 2> $OUT/bench_cfg1.err; timeout 20 python tools/bench_brief.py $OUT/bench_r02_cfg1.json cfg1
timeout -s KILL 240 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file $OUT/r02_step_cfg2.csv python tools/profile_step.py cfg2 > $OUT/ncu_cfg2.log 2>&1
echo "ncu rc=$?"
timeout 60 python tools/ncu_step_summary.py $OUT/r02_step_cfg2.csv > $OUT/r02_step_cfg2.txt 2>/dev/null; head -12 $OUT/r02_step_cfg2.txt | cut -c1-170
