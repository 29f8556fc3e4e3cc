#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_policy_gpu.py tests/test_update_gpu.py -x -q > gpurun_out/r03d_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r03d_tests.log
for v in default tokm_r1 tokm_r2w4 tokm_r3w6 tokm_r4w4; do
  unset MARL_MAZE_LIB
  if [ $v != default ]; then export MARL_MAZE_LIB=$PWD/variants/$v.so; fi
  timeout 300 python tools/tok_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r03d_tok.jsonl
done
