#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 3 -c 1 -o gpurun_out/r02_tcconv18_lean python scripts/prof_kernel.py 64 64 64 18 18 3 > gpurun_out/ncu_d1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_wgrad2_kernel -s 3 -c 1 -o gpurun_out/r02_tcwgrad18_lean python scripts/prof_kernel.py 64 64 64 18 18 3 > gpurun_out/ncu_d2.log 2>&1
ls -la gpurun_out/*.ncu-rep
