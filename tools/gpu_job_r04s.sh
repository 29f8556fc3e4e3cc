#!/bin/bash
# token kernel v3 with compiler-scheduled (non-volatile) HMMAs
mkdir -p gpurun_out
for v in default tokp tokp_h0_w12 tokp_h0_w16b1 tokp tokp_h0_w12; do
  unset MARL_MAZE_LIB
  if [ $v != default ]; then export MARL_MAZE_LIB=$PWD/variants/$v.so; fi
  timeout 300 python tools/tok_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r04s_tok.jsonl
done
for v in tokp tokp_h0_w12; do
export MARL_MAZE_LIB=$PWD/variants/$v.so
timeout 600 python -m pytest tests/test_policy_gpu.py -x -q 2>&1 | tail -2
timeout 300 python tools/k4_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r04s_k4.jsonl
done
