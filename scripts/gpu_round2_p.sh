#!/bin/bash
mkdir -p gpurun_out
for s in "64 64 64 18 18 3" "64 32 32 36 36 3" "64 16 16 72 72 3"; do for st in 4 1; do
  echo "== $s HCM_TC_ST=$st"
  HCM_TC_ST=$st timeout 120 python scripts/prof_kernel.py $s 2>&1 | grep "tc_conv " | sed 's/tc_wgrad.*//'
  HCM_TC_ST=$st HCM_TC_DEBUG=1 timeout 120 python scripts/prof_kernel.py $s > gpurun_out/tmp_p.txt 2>&1
  grep "tc_conv time" gpurun_out/tmp_p.txt | tail -n 2;  grep "tc_conv dbg" gpurun_out/tmp_p.txt | tail -n 1
done; done
