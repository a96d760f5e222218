#!/bin/bash
mkdir -p gpurun_out
N=${NGPU:-8}
HCM_BENCH_DIAG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 8 --warmup 3 --no-extra --no-gpu-eager --no-cpu-baseline > gpurun_out/bench_ad.json 2> gpurun_out/bench_ad.err; echo "rc $?"
grep "\[diag\]" gpurun_out/bench_ad.err | cut -c1-330
python -c "
import json
d=json.loads(open('gpurun_out/bench_ad.json').read().strip().splitlines()[-1]);print(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],d.get('replicas_identical'))"
