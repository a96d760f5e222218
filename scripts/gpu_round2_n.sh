#!/bin/bash
mkdir -p gpurun_out
for s in "64 64 64 18 18 3" "64 32 32 36 36 3"; do
for m in 0 1 2 4 3 5 7; do
  echo "== $s mode $m"
  HCM_TC_DBGMODE=$m timeout 120 python scripts/prof_kernel.py $s 2>&1 | grep "tc_conv " | sed 's/tc_wgrad.*//'
  HCM_TC_DBGMODE=$m HCM_TC_DEBUG=1 timeout 120 python scripts/prof_kernel.py $s 2>&1 | grep "tc_conv dbg" | tail -n 1
done; done > gpurun_out/dbg_n.txt 2>&1
cat gpurun_out/dbg_n.txt
