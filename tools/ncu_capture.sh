#!/bin/bash
# Round-2 ncu evidence, run under gpurun (one GPU): every kernel of one step of the bench workloads with a light metric
# list (device time, DRAM bytes, DRAM / tensor / shared-memory pipe utilisation), plus `--set full` captures of single
# launches of the hot kernels. Each command runs plain first (the recipe's rule), then under ncu. CSV / text only comes back.
set -u
OUT=gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,launch__block_size
step() {   # name, then the profile_step.py arguments
  local name=$1; shift
  timeout 300 env "$@" python tools/profile_step.py $ARGS > $OUT/plain_$name.log 2>&1 &&
  timeout 900 env "$@" ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file $OUT/r02_step_$name.csv \
      python tools/profile_step.py $ARGS > $OUT/ncu_$name.log 2>&1
  echo "step $name rc=$?"
  python tools/ncu_step_summary.py $OUT/r02_step_$name.csv > $OUT/r02_step_$name.txt 2>/dev/null
}
full() {   # name, kernel regex, launch to capture (skip count)
  local name=$1 re=$2 skip=${3:-0}
  timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:$re -s $skip -c 1 -o $OUT/r02_full_$name -f \
      python tools/profile_step.py $ARGS > $OUT/ncu_full_$name.log 2>&1
  echo "full $name rc=$?"
  ncu -i $OUT/r02_full_$name.ncu-rep --page details > $OUT/r02_full_$name.txt 2>/dev/null
  rm -f $OUT/r02_full_$name.ncu-rep
}
ARGS="cfg2";        step cfg2 X=1
ARGS="cfg2";        full lstm_tcw_kernel lstm_tcw_kernel
ARGS="cfg2";        full tc_conv_128_2 "tc_conv_kernel<128,.2" 3
ARGS="cfg2";        full tc_conv_64_2 "tc_conv_kernel<64,.2" 1
ARGS="cfg3 8";      step cfg3 X=1
ARGS="cfg3 8";      full tc_conv_32_2_gn_on_load "tc_conv_kernel<32,.2" 0
ARGS="cfg1";        step cfg1_ecdc PROFILE_ECDC=1
ls -la $OUT | grep r02_
