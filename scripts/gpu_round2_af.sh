#!/bin/bash
mkdir -p gpurun_out
N=${NGPU:-4}
for extra in "" "--no-graph"; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 6 --warmup 3 --no-extra --no-gpu-eager --no-cpu-baseline $extra > gpurun_out/bench_af.json 2> gpurun_out/bench_af.err; echo "rc $? ($extra)"
python -c "
import json
d=json.loads(open('gpurun_out/bench_af.json').read().strip().splitlines()[-1]);print(d['n_gpus'],d['value'],d['ms_per_step'],d.get('replicas_identical'),d.get('replicas_identical_detail'))"
done
