#!/bin/bash
# fused trunk: every other cluster starts late (layer phases of neighbouring clusters interleave in the L2)
mkdir -p gpurun_out
for v in default tf_stag12k tf_stag24k default tf_stag12k tf_stag24k; do
  unset MARL_MAZE_LIB
  if [ $v != default ]; then export MARL_MAZE_LIB=$PWD/variants/$v.so; fi
  timeout 300 python tools/k4_bench.py --no-value 2>&1 | tail -1 | tee -a gpurun_out/r04f_k4.jsonl
done
for v in tf_stag24k_prof; do
  echo $v | tee -a gpurun_out/r04f_trunk_profile.jsonl
  MARL_MAZE_LIB=$PWD/variants/$v.so timeout 300 python tools/trunk_profile.py | tee -a gpurun_out/r04f_trunk_profile.jsonl
done
