#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "tc_ or batchnorm" > gpurun_out/pytest_x.log 2>&1; echo "kernel tests rc $?"; tail -n 3 gpurun_out/pytest_x.log
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "step_matches_oracle and tensorcore or graph_replay" > gpurun_out/pytest_x2.log 2>&1; echo "parity rc $?"; tail -n 3 gpurun_out/pytest_x2.log
for pdl in 1 0; do
HCM_PDL=$pdl timeout 900 python bench.py --steps 10 --warmup 3 --no-extra --no-gpu-eager --no-cpu-baseline > gpurun_out/bench_x$pdl.json 2> gpurun_out/bench_x$pdl.err; echo "bench pdl=$pdl rc $?"
done
python -c "
import json
for f in ('bench_x1','bench_x0'):
    d=json.load(open('gpurun_out/%s.json'%f));print(f,d['value'],d['ms_per_step'],d['e2e']['value'], d['roofline']['us_per_launch'], d['roofline']['frac'])"
tail -n 5 gpurun_out/bench_x1.err
