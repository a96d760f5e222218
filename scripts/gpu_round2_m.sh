#!/bin/bash
mkdir -p gpurun_out
for s in "64 64 64 18 18 3" "64 32 32 36 36 3" "64 16 16 72 72 3" "64 8 8 144 144 3" "64 64 64 64 64 3" "64 64 64 64 256 1"; do
  echo "== $s"
  HCM_TC_DEBUG=1 timeout 120 python scripts/prof_kernel.py $s > gpurun_out/tmp_m.txt 2>&1
  grep "tc_conv dbg" gpurun_out/tmp_m.txt | tail -n 1
  grep "tc_wgrad dbg" gpurun_out/tmp_m.txt | tail -n 1
done > gpurun_out/dbg_m.txt 2>&1
cat gpurun_out/dbg_m.txt
