#!/bin/bash
# token kernel third generation with the exact fp32 residual
mkdir -p gpurun_out
for i in 1 2; do timeout 300 python tools/tok_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r04v_tok.jsonl; done
timeout 900 python -m pytest tests/test_policy_gpu.py tests/test_ppo_gpu.py tests/test_update_gpu.py -q 2>&1 | tail -4
timeout 300 python tools/k4_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r04v_k4.jsonl
