#!/bin/bash
# K2 with L2 cache hints (1 = state evict_last, 2 = streams evict_first, 3 = both) against the default
mkdir -p gpurun_out
for v in default k2_l2hint1 k2_l2hint2 k2_l2hint3 default k2_l2hint1 k2_l2hint2 k2_l2hint3; do
  unset MARL_MAZE_LIB
  if [ $v != default ]; then export MARL_MAZE_LIB=$PWD/variants/$v.so; fi
  timeout 300 python tools/k2_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r04m_k2.jsonl
done
for v in k2_l2hint3; do
  MARL_MAZE_LIB=$PWD/variants/$v.so timeout 300 python tools/k2_bench.py --envs 524288 2>&1 | tail -1 | tee -a gpurun_out/r04m_k2.jsonl
  MARL_MAZE_LIB=$PWD/variants/$v.so timeout 600 python -m pytest tests/test_env_parity_gpu.py -x -q 2>&1 | tail -2
done
unset MARL_MAZE_LIB
timeout 300 python tools/k2_bench.py --envs 524288 2>&1 | tail -1 | tee -a gpurun_out/r04m_k2.jsonl
