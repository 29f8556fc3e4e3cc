#!/bin/bash
# r02p: uniform-register MMA / TMA issue in every tcgen05 kernel: full GPU tests, GEMM kernel timings, rollout + update
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02p_tests.log 2>&1; echo "all tests rc=$?"; tail -8 gpurun_out/r02p_tests.log
timeout 300 python tools/k5_bench.py --skip-update 2>&1 | tail -1 | tee gpurun_out/r02p_k5_kernels.json
timeout 300 python tools/k4_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r02p_k4.jsonl
timeout 300 python tools/k4_bench.py --no-fused 2>&1 | tail -1 | tee -a gpurun_out/r02p_k4.jsonl
timeout 600 python tools/rollout_bench.py --epochs 3 2>&1 | tail -1 | tee gpurun_out/r02p_rollout.json
