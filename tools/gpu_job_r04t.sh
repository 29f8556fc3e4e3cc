#!/bin/bash
# token kernel v3 with the next row's observation columns prefetched
mkdir -p gpurun_out
for v in tokp tokp_w10 tokp_w12 tokp tokp_w10 tokp_w12; do
  export MARL_MAZE_LIB=$PWD/variants/$v.so
  timeout 300 python tools/tok_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r04t_tok.jsonl
done
export MARL_MAZE_LIB=$PWD/variants/tokp.so
timeout 600 python -m pytest tests/test_policy_gpu.py -x -q 2>&1 | tail -2
