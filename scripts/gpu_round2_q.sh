#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "tc_" > gpurun_out/pytest_q.log 2>&1; echo "kernel tests rc $?"; tail -n 3 gpurun_out/pytest_q.log
for st in 4 1; do HCM_TC_ST=$st timeout 300 python scripts/time_shapes.py "64 64 64 18 18 3" "32 64 64 18 18 3" "16 96 96 32 32 3" ; done 2>&1 | tee gpurun_out/time_q.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-extra --no-gpu-eager --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench rc $?"
HCM_TC_ST=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-extra --no-gpu-eager --no-cpu-baseline > gpurun_out/bench_q_st1.json 2> gpurun_out/bench_q_st1.err; echo "bench rc $?"
python -c "
import json
for f in ('bench_q','bench_q_st1'):
    d=json.load(open('gpurun_out/%s.json'%f));print(f,d['value'],d['ms_per_step'],d['e2e']['value'], d['roofline']['us_per_launch'], d['roofline']['frac'])"
