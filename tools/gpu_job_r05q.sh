#!/bin/bash
mkdir -p gpurun_out
for v in default tokp_w4b4 tokp_w16b1 tokp_w5b3 default tokp_w4b4; do
  unset MARL_MAZE_LIB
  if [ $v != default ]; then export MARL_MAZE_LIB=$PWD/variants/$v.so; fi
  timeout 300 python tools/tok_bench.py 2>&1 | tail -1 | cut -c1-120 | tee -a gpurun_out/r05q_tok.jsonl
done
