#!/bin/bash
# fused trunk with dedicated layer-0 landing slots (weight ring 6 or 7 stages) against the default (8 stages, slots alias the lo plane)
mkdir -p gpurun_out
for v in tf_ad_w6a2 tf_ad_w7a1; do
  MARL_MAZE_LIB=$PWD/variants/$v.so timeout 600 python -m pytest tests/test_policy_gpu.py -x -q -k fused 2>&1 | tail -2
done
for v in default tf_ad_w6a2 tf_ad_w7a1 default tf_ad_w6a2 tf_ad_w7a1; do
  unset MARL_MAZE_LIB
  if [ $v != default ]; then export MARL_MAZE_LIB=$PWD/variants/$v.so; fi
  timeout 300 python tools/k4_bench.py --no-value 2>&1 | tail -1 | tee -a gpurun_out/r04e_k4.jsonl
done
for v in tf_prof tf_ad_w6a2_prof; do
  echo $v | tee -a gpurun_out/r04e_trunk_profile.jsonl
  MARL_MAZE_LIB=$PWD/variants/$v.so timeout 300 python tools/trunk_profile.py | tee -a gpurun_out/r04e_trunk_profile.jsonl
done
