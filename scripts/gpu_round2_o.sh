#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "tc_ or dense" > gpurun_out/pytest_o.log 2>&1; echo "kernel tests rc $?"; tail -n 3 gpurun_out/pytest_o.log
timeout 300 python scripts/time_shapes.py "64 64 64 18 18 3" "64 32 32 36 36 3" "64 16 16 72 72 3" "64 8 8 144 144 3" "64 64 64 64 64 3" "64 64 64 64 256 1" "64 64 64 256 64 1" > gpurun_out/time_o.txt 2>&1; cat gpurun_out/time_o.txt
for s in "64 64 64 18 18 3" "64 32 32 36 36 3"; do for m in 0 7; do
  HCM_TC_DBGMODE=$m HCM_TC_DEBUG=1 timeout 120 python scripts/prof_kernel.py $s > gpurun_out/tmp_o.txt 2>&1
  grep "tc_conv dbg" gpurun_out/tmp_o.txt | tail -n 1;  grep "tc_wgrad dbg" gpurun_out/tmp_o.txt | tail -n 1
done; done
