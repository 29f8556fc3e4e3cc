#!/bin/bash
# token kernel, third generation (keys / queries / values as tensor-path products of the token tile) against the second
mkdir -p gpurun_out
for v in default tokp tokp_w10 tokp_w6b3; do
  unset MARL_MAZE_LIB
  if [ $v != default ]; then export MARL_MAZE_LIB=$PWD/variants/$v.so; fi
  timeout 300 python tools/tok_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r04n_tok.jsonl
done
export MARL_MAZE_LIB=$PWD/variants/tokp.so
timeout 600 python -m pytest tests/test_policy_gpu.py -x -q 2>&1 | tail -3
for v in default tokp default tokp; do
  unset MARL_MAZE_LIB
  if [ $v != default ]; then export MARL_MAZE_LIB=$PWD/variants/$v.so; fi
  timeout 300 python tools/k4_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r04n_k4.jsonl
done
