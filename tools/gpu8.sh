#!/bin/bash
# 8-GPU pass: host D2H ceiling, bench at N=8 (and N=1 on the same box for the efficiency), C4 sweep at N=8
mkdir -p gpurun_out
N=${1:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/d2h_ceiling.py > gpurun_out/r2_d2h_ceiling_n$N.log 2>&1; grep -E "^\{" gpurun_out/r2_d2h_ceiling_n$N.log
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_n1_samebox.json 2> gpurun_out/r2_bench_n1_samebox.err; python -c "
import json;d=json.load(open('gpurun_out/r2_bench_n1_samebox.json'));print('N=1 value',round(d['value']),'e2e',round(d['e2e']['value']))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; python -c "
import json;d=json.load(open('gpurun_out/r2_bench_n$N.json'));print('N=$N value',round(d['value']),'e2e',round(d['e2e']['value']), d['e2e'].get('ms_per_step_runs'))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 -m robosimgs_b200.sweep > gpurun_out/r2_c4_sweep_n$N.json 2> gpurun_out/r2_c4_sweep_n$N.err; tail -c 400 gpurun_out/r2_c4_sweep_n$N.json
