#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_policy_gpu.py tests/test_update_gpu.py -x -q > gpurun_out/r02s_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02s_tests.log
for v in default tokm_libm tokm_w8 tokm_w6 tokm_w2; do
  unset MARL_MAZE_LIB
  if [ $v != default ]; then export MARL_MAZE_LIB=$PWD/variants/$v.so; fi
  timeout 300 python tools/tok_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r02s_tok.jsonl
done
