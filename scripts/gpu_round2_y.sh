#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "fuse or adjoint or seg_head or tc_conv" > gpurun_out/pytest_y.log 2>&1; echo "kernel tests rc $?"; tail -n 3 gpurun_out/pytest_y.log
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "step_matches_oracle and tensorcore or seg_" > gpurun_out/pytest_y2.log 2>&1; echo "parity rc $?"; tail -n 3 gpurun_out/pytest_y2.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-extra --no-gpu-eager --no-cpu-baseline > gpurun_out/bench_y.json 2> gpurun_out/bench_y.err; echo "bench rc $?"
timeout 900 python bench.py --steps 10 --warmup 3 --no-extra --no-gpu-eager --no-cpu-baseline --stage 2 --batch 32 > gpurun_out/bench_y2.json 2> gpurun_out/bench_y2.err; echo "bench rc $?"
python -c "
import json
for f in ('bench_y','bench_y2'):
    d=json.load(open('gpurun_out/%s.json'%f));print(f,d['value'],d['ms_per_step'],d['e2e']['value']); k=d['kernel_families_ms']; c=d['kernel_families_calls']; print({n:(round(k[n],2),c[n]) for n in ('fuse_sum','upsample_adjoint','gemm')})"
