#!/bin/bash
# `--set full` captures of single tc_conv launches (ncu's -k matches the function name without template arguments, so the
# instance is picked by its position among the tc_conv launches of the step): config 2: launch 6 = down128 <128,2,6>,
# launch 0 = down32 <64,2,4>; config 3: launch 0 = the first block conv with GroupNorm on load <32,2,4>.
set -u
OUT=gpurun_out
full() {   # name, skip, profile_step args...
  local name=$1 skip=$2; shift; shift
  timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:tc_conv_kernel -s $skip -c 1 -o $OUT/r02_full_$name -f \
      python tools/profile_step.py "$@" > $OUT/ncu_full_$name.log 2>&1
  echo "full $name rc=$?"
  ncu -i $OUT/r02_full_$name.ncu-rep --page details > $OUT/r02_full_$name.txt 2>/dev/null
  rm -f $OUT/r02_full_$name.ncu-rep
}
full tc_conv_128_2_6_down128 6 cfg2
full tc_conv_64_2_4_down32 0 cfg2
full tc_conv_32_2_4_gn_on_load 0 cfg3 8
ls -la $OUT | grep r02_full
