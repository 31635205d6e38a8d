#!/bin/bash
# End-of-round measurement pass (one GPU): tests, smoke, bench (both arms), launch list, ncu full capture
# of the main kernels, naive comparator, all BASELINE configs.  Outputs under gpurun_out/final_*.
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; tail -2 gpurun_out/final_pytest.log
timeout 200 python __graft_entry__.py --smoke > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_target.py > gpurun_out/final_san_memcheck.log 2>&1; grep -E "ERROR SUMMARY" gpurun_out/final_san_memcheck.log; grep -c "^ok" gpurun_out/final_san_memcheck.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/final_clocks.csv &
SMI=$!
timeout 400 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 600 gpurun_out/final_bench.json
kill $SMI
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; tail -c 300 gpurun_out/final_bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final_launches_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_render|k_project|k_emit|k_tile_ranges|k_fix|Onesweep|k_photometric" -s 22 -c 22 -o gpurun_out/final_prof python tools/ncu_target.py 3 > gpurun_out/final_ncu.log 2>&1; tail -1 gpurun_out/final_ncu.log
timeout 300 python tools/compare_naive.py > gpurun_out/final_naive.log 2>&1; tail -c 400 gpurun_out/final_naive.log
timeout 400 python tests/devtools/run_configs.py > gpurun_out/final_configs.log 2>&1; tail -3 gpurun_out/final_configs.log
