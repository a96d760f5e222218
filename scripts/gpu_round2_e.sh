#!/bin/bash
mkdir -p gpurun_out
HCM_LAUNCH_LOG=gpurun_out/launchlog_b64.txt ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_b64.csv python bench.py --ncu-step --no-graph --warmup 3 > gpurun_out/ncu_e.log 2>&1
echo "rc $?"; wc -l gpurun_out/launches_b64.csv gpurun_out/launchlog_b64.txt
