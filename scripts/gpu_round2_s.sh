#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pointnet2_gpu.py -m gpu -x -q -s > gpurun_out/pytest_s.log 2>&1; echo "pn2 tests rc $?"; grep -n "fps B=\|passed\|failed\|Error\|assert" gpurun_out/pytest_s.log | tail -n 25
