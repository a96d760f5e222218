#!/bin/bash
mkdir -p gpurun_out
N=${NGPU:-8}
for sync in 1 0; do
echo "=== SYNC=$sync"
SYNC=$sync BATCH=64 RES=256 NDATA=165894 NCEK=16384 NSTEPS=8 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$sync scripts/check_replicas.py 2>&1 | grep -v "^W\|^\[W\|NCCL\|^\*\|OMP_NUM" | tail -n 30
done
