#!/bin/bash
# Quick GPU pass (one GPU): tests, bench, optional ncu.  usage: gpu_quick.sh TAG [ncu-kernel-regex]
TAG=${1:-q}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_bench.json'))
    print('value',round(d['value']),'single',round(d['single_stream']['ms_per_frame'],4),'train',round(d['train']['ms_per_iter'],4),'e2e',round(d['e2e']['value']))
    print(d['roofline']['stage_ms_train']); print(d.get('parity'))
except Exception as e: print('bench parse failed',e)
PY
if [ -n "$2" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$2" -s ${NCU_SKIP:-22} -c ${NCU_COUNT:-12} -o gpurun_out/${TAG}_prof -f python tools/ncu_target.py 3 > gpurun_out/${TAG}_ncu.log 2>&1; tail -1 gpurun_out/${TAG}_ncu.log
fi
