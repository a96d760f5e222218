#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python bench.py > gpurun_out/bench_z.json 2> gpurun_out/bench_z.err ) 2>&1 | tail -n 4; echo "bench rc $?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_z.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value']); print(json.dumps(d.get('widened_rows'))[:1500])"
tail -n 3 gpurun_out/bench_z.err
