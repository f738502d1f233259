set -u
OUT=gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,launch__block_size
step() {
  local name=$1; shift
  timeout 300 env "$@" python tools/profile_step.py $ARGS > $OUT/plain_$name.log 2>&1 &&
  timeout 900 env "$@" ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file $OUT/r02_step_$name.csv \
      python tools/profile_step.py $ARGS > $OUT/ncu_$name.log 2>&1
  echo "step $name rc=$?"
}
ARGS="cfg3 8";      step cfg3 X=1
ARGS="cfg2";        step cfg2 X=1
