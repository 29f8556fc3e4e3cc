#!/bin/bash
# r02n: timing experiments on the fused trunk kernel: which input stream holds the main loops back
mkdir -p gpurun_out
for v in tf_exp_now tf_exp_noa tf_exp_none; do
  echo "== $v" | tee -a gpurun_out/r02n_exp.jsonl
  MARL_MAZE_LIB=$PWD/variants/$v.so timeout 300 python tools/trunk_profile.py 2>&1 | tail -1 | tee -a gpurun_out/r02n_exp.jsonl
done
