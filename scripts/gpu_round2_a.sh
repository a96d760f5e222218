#!/bin/bash
# first GPU call of round 2: box facts, the full GPU test-suite (incl. BASELINE-shape parity), the bench line with the new legs
mkdir -p gpurun_out
{ nproc; free -g; nvidia-smi -L; nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv; } > gpurun_out/boxinfo.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/boxinfo.txt
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 --detail gpurun_out/detail_b64.txt > gpurun_out/bench_b64.json 2> gpurun_out/bench_b64.err; echo "bench rc $?" >> gpurun_out/boxinfo.txt
tail -c 1500 gpurun_out/bench_b64.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
