#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "tc_" > gpurun_out/pytest_v.log 2>&1; echo "kernel tests rc $?"; tail -n 3 gpurun_out/pytest_v.log
timeout 300 python scripts/time_acc.py "64 64 64 18 18 3" "64 32 32 36 36 3" "64 16 16 72 72 3" "64 8 8 144 144 3" "64 64 64 64 256 1" "64 64 64 256 64 1" "64 64 64 64 64 3" 2>&1 | tee gpurun_out/time_acc2.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-extra --no-gpu-eager --no-cpu-baseline > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err; echo "bench rc $?"
timeout 900 python bench.py --steps 10 --warmup 3 --no-extra --no-gpu-eager --no-cpu-baseline --stage 2 --batch 32 > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; echo "bench rc $?"
python -c "
import json
for f in ('bench_v','bench_v2'):
    d=json.load(open('gpurun_out/%s.json'%f));print(f,d['value'],d['ms_per_step'],d['e2e']['value'], d['roofline']['us_per_launch'], d['roofline']['frac'])"
