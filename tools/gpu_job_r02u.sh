#!/bin/bash
# r02u: incremental pool refill (only consumed slots rebuilt) + in-place background prefetch: tests, rollout timing with / without prefetch
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02u_tests.log 2>&1; echo "all tests rc=$?"; tail -8 gpurun_out/r02u_tests.log
timeout 600 python tools/rollout_bench.py --epochs 4 2>&1 | tail -1 | tee gpurun_out/r02u_rollout.json | cut -c1-600
timeout 600 python tools/rollout_bench.py --epochs 4 --prefetch 2>&1 | tail -1 | tee gpurun_out/r02u_rollout_prefetch.json | cut -c1-600
