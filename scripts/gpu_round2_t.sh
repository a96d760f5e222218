#!/bin/bash
# round-2 evidence run: full GPU test-suite, the default bench line (all legs), the reference arm, the ncu launch list of one
# step and ncu --set full captures of the dominant kernels
mkdir -p gpurun_out
{ nproc; free -g | head -n 2; nvidia-smi -L; nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv; } > gpurun_out/boxinfo_t.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_t.log 2>&1; echo "pytest rc $?" | tee -a gpurun_out/boxinfo_t.txt
tail -n 4 gpurun_out/pytest_gpu_t.log
timeout 900 python bench.py --steps 20 --warmup 5 --detail gpurun_out/r02_detail_b64.txt > gpurun_out/r02_bench_b64.json 2> gpurun_out/r02_bench_b64.err; echo "bench rc $?" | tee -a gpurun_out/boxinfo_t.txt
tail -c 600 gpurun_out/r02_bench_b64.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_cpu.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc $?"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_b64.csv python bench.py --ncu-step --no-graph --warmup 3 > gpurun_out/ncu_t.log 2>&1; echo "ncu list rc $?"
ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 3 -c 1 -o gpurun_out/r02_tcconv18_b64_st4 python scripts/prof_kernel.py 64 64 64 18 18 3 > gpurun_out/ncu_t1.log 2>&1; echo "ncu conv rc $?"
ncu --set full --clock-control none --import-source on -k regex:tc_wgrad2_kernel -s 3 -c 1 -o gpurun_out/r02_tcwgrad18_b64_v5 python scripts/prof_kernel.py 64 64 64 18 18 3 > gpurun_out/ncu_t2.log 2>&1; echo "ncu wgrad rc $?"
ls -la gpurun_out/*.ncu-rep | tail -n 4
