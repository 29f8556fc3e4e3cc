#!/bin/bash
# last check of the committed state: smoke, the full GPU suite, bench.py with its default arguments
mkdir -p gpurun_out
( time timeout 600 python __graft_entry__.py smoke ) > gpurun_out/r05p_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r05p_smoke.log | head -1 | cut -c1-250
( time timeout 1500 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/r05p_pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed" gpurun_out/r05p_pytest_gpu.log
( time timeout 900 python bench.py ) > gpurun_out/r05p_bench_default.json 2> gpurun_out/r05p_bench_default.err; echo "bench rc=$?"; tail -3 gpurun_out/r05p_bench_default.err
