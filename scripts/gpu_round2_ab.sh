#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_ncu.log 2>&1; echo "rc $?"
tail -n 3 gpurun_out/smoke_ncu.log
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/smoke_launches.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
ci={h:i for i,h in enumerate(rows[hi])}
c=collections.Counter()
first={}
for k,r in enumerate(rows[hi+1:]):
    if len(r)<len(ci): continue
    n=r[ci["Kernel Name"]].split("(")[0][:60]
    c[n]+=1; first.setdefault(n,k)
for n,v in c.most_common(25): print(v, first[n], n)
PY
