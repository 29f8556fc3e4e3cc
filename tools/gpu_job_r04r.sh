#!/bin/bash
# token kernel, third generation with ex2.approx softmax and the residual on the tensor path; fragment hoisting / warps per block variants
mkdir -p gpurun_out
for v in tokp_rn tokp_rz tokp_rz_h0 tokp_rz_h0_w12 tokp_rn tokp_rz; do
  unset MARL_MAZE_LIB
  if [ $v != default ]; then export MARL_MAZE_LIB=$PWD/variants/$v.so; fi
  timeout 300 python tools/tok_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r04r_tok.jsonl
done
export MARL_MAZE_LIB=$PWD/variants/tokp_rz.so
timeout 600 python -m pytest tests/test_policy_gpu.py -x -q 2>&1 | tail -3
