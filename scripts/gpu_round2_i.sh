#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "tc_ or dense" > gpurun_out/pytest_i.log 2>&1; echo "kernel tests rc $?"; tail -5 gpurun_out/pytest_i.log
timeout 300 python scripts/time_shapes.py "64 16 16 72 72 3" "64 8 8 144 144 3" "16 12 12 256 256 3" "64 64 64 64 64 3" "64 64 64 256 64 1" "64 64 64 18 18 3" > gpurun_out/time_i.txt 2>&1; cat gpurun_out/time_i.txt
HCM_TC_DEBUG=1 python scripts/prof_kernel.py 64 8 8 144 144 3 2>&1 | grep "tc_conv dbg" | tail -1
timeout 120 python scripts/dbg_dense.py 2>&1 | tail -9
timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-gpu-eager --no-cpu-baseline > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err; echo "bench rc $?"
python -c "
import json;d=json.load(open('gpurun_out/bench_i.json'));print(d['value'],d['ms_per_step'],d['e2e']['value']);o=d['roofline_other'];print(o['stage4_conv']['us_per_launch'],o['dense_affinity_fwd']['us_per_launch'],o['dense_affinity_bwd']['us_per_launch'])"
