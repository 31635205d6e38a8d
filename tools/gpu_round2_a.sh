#!/bin/bash
# Round-2 GPU pass A (one GPU): full GPU test-suite incl. the new C3/C4 parity tests, bench with parity block.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2a_pytest.log 2>&1; tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 1500 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
