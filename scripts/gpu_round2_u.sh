#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; echo "2gpu rc $?"
tail -c 900 gpurun_out/r02_bench_2gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_2gpu.json 2> gpurun_out/r02_bench_ref_2gpu.err; echo "ref 2gpu rc $?"; tail -c 300 gpurun_out/r02_bench_ref_2gpu.json
