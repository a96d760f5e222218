#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./scripts/bench_umma2.bin > gpurun_out/umma2.txt 2>&1
HCM_PARITY_SIMT=1 timeout 1500 python -m pytest tests/test_parity_gpu.py -m gpu -q -s -k "baseline_shape or two_ranks" > gpurun_out/pytest_b.log 2>&1; echo "pytest rc $?"
grep -n "grads\|eager\|SIMT\|after SGD\|rank\|passed\|failed" gpurun_out/pytest_b.log | tail -40
