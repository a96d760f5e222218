#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "dense or stage2" > gpurun_out/pytest_k.log 2>&1; echo "kernel tests rc $?"; tail -5 gpurun_out/pytest_k.log
timeout 120 python scripts/dbg_dense.py 2>&1 | tail -9
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "step_matches_oracle and tensorcore" > gpurun_out/pytest_k2.log 2>&1; echo "parity rc $?"; tail -3 gpurun_out/pytest_k2.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-gpu-eager --no-cpu-baseline --stage 2 --batch 32 > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; echo "bench rc $?"
python -c "
import json;d=json.load(open('gpurun_out/bench_k.json'));print(d['value'],d['ms_per_step'],d['e2e']['value']);o=d['roofline_other'];print(o['dense_affinity_fwd']['us_per_launch'],o['dense_affinity_bwd']['us_per_launch'])"
