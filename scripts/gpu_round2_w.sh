#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "padded_stem or tc_" > gpurun_out/pytest_w.log 2>&1; echo "kernel tests rc $?"; tail -n 3 gpurun_out/pytest_w.log
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "step_matches_oracle and tensorcore or golden" > gpurun_out/pytest_w2.log 2>&1; echo "parity rc $?"; tail -n 3 gpurun_out/pytest_w2.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-extra --no-gpu-eager --no-cpu-baseline --detail gpurun_out/detail_w.txt > gpurun_out/bench_w.json 2> gpurun_out/bench_w.err; echo "bench rc $?"
python -c "
import json
d=json.load(open('gpurun_out/bench_w.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'], d['roofline']['us_per_launch'], d['roofline']['frac'])"
head -n 24 gpurun_out/detail_w.txt
