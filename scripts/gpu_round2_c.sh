#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "tc_ or dense" > gpurun_out/pytest_c.log 2>&1; echo "kernel tests rc $?"; tail -3 gpurun_out/pytest_c.log
for shp in "64 64 64 18 18 3" "64 32 32 36 36 3" "64 16 16 72 72 3" "64 8 8 144 144 3" "64 64 64 64 64 3" "64 64 64 64 256 1" "16 96 96 32 32 3" "16 48 48 64 64 3"; do
  echo "== $shp"; HCM_TC_DEBUG=1 timeout 120 python scripts/prof_kernel.py $shp 2>&1 | awk '/dbg\]/{a[$1" "$2]=$0} /TFLOP/{print} END{for(k in a)print a[k]}'
done > gpurun_out/prof_c.txt 2>&1
cat gpurun_out/prof_c.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-gpu-eager --no-cpu-baseline --detail gpurun_out/detail_c.txt > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; echo "bench rc $?"
python -c "
import json;d=json.load(open('gpurun_out/bench_c.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['us_per_launch'],d['roofline']['frac'])"
head -12 gpurun_out/detail_c.txt
