#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_j.log 2>&1; echo "gpu tests rc $?"; tail -4 gpurun_out/pytest_j.log
grep -n "grads vs\|eager-GPU\|after SGD" gpurun_out/pytest_j.log | head -20
ncu --set full --clock-control none --import-source on -k regex:dense_affinity_kernel -s 4 -c 2 -o gpurun_out/r02_dense_v2 python scripts/dbg_dense.py > gpurun_out/ncu_j1.log 2>&1
