#!/bin/bash
# ncu launch list (per-launch device time, cold-cache and serialised: compare SHARES) of a few C3 fwd+bwd iterations
TAG=${1:-l}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python tools/ncu_target.py 4 > gpurun_out/${TAG}_launches.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/${TAG}_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
last=[(r[ki].split('(')[0][:60], float(r[vi].replace(',',''))) for r in rows[1:]]
# print the last iteration (from the last k_project on)
idx=max(i for i,(k,v) in enumerate(last) if k.startswith('k_project<') or k.startswith('void b200gs::k_project<') or 'k_project<' in k)
prev=[i for i,(k,v) in enumerate(last) if 'k_project<' in k]
start=prev[-1]
for k,v in last[start:]: print('%9.2f us  %s'%(v/1000.0,k))
PY
