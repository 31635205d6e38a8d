#!/bin/bash
# Round-2 measurement pass (one GPU): tests, smoke, sanitizers, bench (both arms), launch list of the bench command,
# ncu full capture of the main kernels, BASELINE configs.  Outputs under gpurun_out/r2_*.
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -2 gpurun_out/${TAG}_pytest.log
timeout 200 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize_target.py quick > gpurun_out/${TAG}_san_${tool}.log 2>&1
  echo "$tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${TAG}_san_${tool}.log | tail -1) ok-lines=$(grep -c '^ok' gpurun_out/${TAG}_san_${tool}.log)"
done
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/${TAG}_clocks.csv &
SMI=$!
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 400 gpurun_out/${TAG}_bench.json; echo
kill $SMI
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; tail -c 200 gpurun_out/${TAG}_bench_reference.json; echo
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 3 --repeats 1 --no-cpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_render|k_project|k_emit|k_bucket|k_photometric" -s 27 -c 27 -o gpurun_out/${TAG}_prof -f python tools/ncu_target.py 4 > gpurun_out/${TAG}_ncu.log 2>&1; tail -1 gpurun_out/${TAG}_ncu.log
timeout 600 python tests/devtools/run_configs.py > gpurun_out/${TAG}_configs.log 2>&1; tail -3 gpurun_out/${TAG}_configs.log
