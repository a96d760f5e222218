#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "tc_" > gpurun_out/pytest_f.log 2>&1; echo "kernel tests rc $?"; tail -3 gpurun_out/pytest_f.log
timeout 300 python scripts/time_shapes.py "64 64 64 18 18 3" "64 32 32 36 36 3" "64 16 16 72 72 3" "64 8 8 144 144 3" "32 8 8 144 144 3" "64 64 64 64 64 3" "64 64 64 64 256 1" "64 64 64 256 64 1" "64 16 16 72 18 1" "16 96 96 32 32 3" "16 12 12 256 256 3" > gpurun_out/time_f.txt 2>&1; cat gpurun_out/time_f.txt
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "step_matches_oracle and tensorcore" > gpurun_out/pytest_f2.log 2>&1; echo "parity rc $?"; tail -3 gpurun_out/pytest_f2.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-gpu-eager --no-cpu-baseline > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; echo "bench rc $?"
python -c "
import json;d=json.load(open('gpurun_out/bench_f.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline_other']['stage4_conv'])"
