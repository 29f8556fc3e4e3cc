#!/bin/bash
mkdir -p gpurun_out
MARL_MAZE_LIB=$PWD/variants/tf_e16.so timeout 600 python -m pytest tests/test_policy_gpu.py -x -q 2>&1 | tail -2
MARL_MAZE_LIB=$PWD/variants/tf_e16_prof.so timeout 300 python tools/trunk_profile.py | tee -a gpurun_out/r03i_trunk_profile.json
for v in default tf_e16 default tf_e16; do
  unset MARL_MAZE_LIB
  if [ $v != default ]; then export MARL_MAZE_LIB=$PWD/variants/$v.so; fi
  timeout 300 python tools/k4_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r03i_k4.jsonl
done
