#!/bin/bash
# Last measurement pass of round 2 (one GPU, ~6 min of box time): tests, smoke, memcheck, bench (both arms), launch list of
# the bench command, ncu full capture of the main kernels.  Outputs under gpurun_out/${TAG}_*.
TAG=${1:-r2c}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -2 gpurun_out/${TAG}_pytest.log
timeout 200 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_target.py quick > gpurun_out/${TAG}_san_memcheck.log 2>&1
echo "memcheck: $(grep -E 'ERROR SUMMARY' gpurun_out/${TAG}_san_memcheck.log | tail -1) ok-lines=$(grep -c '^ok' gpurun_out/${TAG}_san_memcheck.log)"
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/${TAG}_clocks.csv &
SMI=$!
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 300 gpurun_out/${TAG}_bench.json; echo
kill $SMI
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; tail -c 200 gpurun_out/${TAG}_bench_reference.json; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 3 --repeats 1 --no-cpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_render|k_project|k_emit|k_bucket|k_photometric" -s 27 -c 27 -o gpurun_out/${TAG}_prof -f python tools/ncu_target.py 4 > gpurun_out/${TAG}_ncu.log 2>&1; tail -1 gpurun_out/${TAG}_ncu.log
