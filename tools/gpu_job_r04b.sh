#!/bin/bash
# critic v2 (lane = environment, 64 neurons in registers) against v1: parity tests, stand-alone timing, policy forward, rollout
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_policy_gpu.py tests/test_ppo_gpu.py -x -q 2>&1 | tail -3
for v in default critic_w12 critic_w8 default critic_w12 critic_w8; do
  unset MARL_MAZE_LIB
  if [ $v != default ]; then export MARL_MAZE_LIB=$PWD/variants/$v.so; fi
  timeout 300 python tools/critic_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r04b_critic.jsonl
  timeout 300 python tools/critic_bench.py --envs 1048576 2>&1 | tail -1 | tee -a gpurun_out/r04b_critic.jsonl
  timeout 300 python tools/k4_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r04b_k4.jsonl
done
unset MARL_MAZE_LIB
timeout 600 python tools/rollout_bench.py --epochs 3 2>&1 | tail -3 | tee gpurun_out/r04b_rollout.json
