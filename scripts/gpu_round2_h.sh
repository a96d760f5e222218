#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:dense_affinity_kernel -s 4 -c 2 -o gpurun_out/r02_dense python scripts/dbg_dense.py > gpurun_out/ncu_h1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 3 -c 1 -o gpurun_out/r02_tcconv144 python scripts/prof_kernel.py 64 8 8 144 144 3 > gpurun_out/ncu_h2.log 2>&1
HCM_TC_DEBUG=1 python scripts/prof_kernel.py 64 8 8 144 144 3 2>&1 | grep "tc_conv dbg" | tail -1
HCM_TC_DEBUG=1 python scripts/prof_kernel.py 16 12 12 256 256 3 2>&1 | grep "tc_conv dbg" | tail -1
ls -la gpurun_out/*.ncu-rep
