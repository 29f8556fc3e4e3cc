#!/bin/bash
# HMMA critic with the staged prologue; low-latency single-env path (pinned mirrors, get_action through the K4 kernels)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_policy_gpu.py tests/test_ppo_gpu.py -x -q 2>&1 | tail -5
for i in 1 2; do
  timeout 300 python tools/critic_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r04c_critic.jsonl
  timeout 300 python tools/critic_bench.py --envs 1048576 2>&1 | tail -1 | tee -a gpurun_out/r04c_critic.jsonl
  timeout 300 python tools/k4_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r04c_k4.jsonl
done
timeout 300 python tools/single_env_bench.py 2>&1 | tail -1 | tee gpurun_out/r04c_single_env.json
