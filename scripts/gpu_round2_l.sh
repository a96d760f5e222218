#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "stage_input" > gpurun_out/pytest_l.log 2>&1; echo "kernel tests rc $?"; tail -n 15 gpurun_out/pytest_l.log
