#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "seg_head" > gpurun_out/pytest_r.log 2>&1; echo "kernel tests rc $?"; tail -n 12 gpurun_out/pytest_r.log
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -s -k "seg_" > gpurun_out/pytest_r2.log 2>&1; echo "parity rc $?"; grep -n "seg step\|encoder grads\|passed\|failed\|Error\|assert" gpurun_out/pytest_r2.log | tail -n 20
