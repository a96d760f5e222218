#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "batchnorm or stage2 or seg_head" > gpurun_out/pytest_aa.log 2>&1; echo "kernel tests rc $?"; tail -n 3 gpurun_out/pytest_aa.log
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "step_matches_oracle or golden or graph_replay or seg_" > gpurun_out/pytest_aa2.log 2>&1; echo "parity rc $?"; tail -n 3 gpurun_out/pytest_aa2.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-extra --no-gpu-eager --no-cpu-baseline > gpurun_out/bench_aa.json 2> gpurun_out/bench_aa.err; echo "bench rc $?"
python -c "
import json
d=json.load(open('gpurun_out/bench_aa.json'));print(d['value'],d['ms_per_step'],d['e2e']['value']); k=d['kernel_families_ms']; print({n:round(k[n],2) for n in k if n.startswith('bn_')})"
