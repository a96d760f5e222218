#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "two_ranks" > gpurun_out/pytest_ag.log 2>&1; echo "two-rank rc $?"; tail -n 2 gpurun_out/pytest_ag.log
