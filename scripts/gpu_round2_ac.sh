#!/bin/bash
mkdir -p gpurun_out
N=${NGPU:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo "${N}gpu rc $?"
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_${N}gpu.json').read().strip().splitlines()[-1]);print(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],d.get('replicas_identical'),{k:round(v['value'],1) for k,v in d.get('extra_configs',{}).items()})"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_${N}gpu.json 2> gpurun_out/r02_bench_ref_${N}gpu.err; echo "ref rc $?"; tail -c 200 gpurun_out/r02_bench_ref_${N}gpu.json
