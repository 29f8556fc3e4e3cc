#!/bin/bash
# token kernel third generation as the default: policy / PPO / update tests, policy forward and rollout against the second generation, ncu of the kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_policy_gpu.py tests/test_ppo_gpu.py tests/test_update_gpu.py -x -q 2>&1 | tail -3
for v in default tok_gen2 default tok_gen2; do
  unset MARL_MAZE_LIB
  if [ $v != default ]; then export MARL_MAZE_LIB=$PWD/variants/$v.so; fi
  timeout 300 python tools/k4_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r04u_k4.jsonl
done
unset MARL_MAZE_LIB
timeout 600 python tools/rollout_bench.py --epochs 3 2>&1 | tail -1 | tee gpurun_out/r04u_rollout.json
bash tools/gpu_ncu_tok.sh r04u
