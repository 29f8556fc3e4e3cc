#!/bin/bash
# r02q: token kernel on the warp-level tensor path (mm_tokens_mma.cu): parity (policy + update tests), timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_policy_gpu.py tests/test_update_gpu.py -x -q > gpurun_out/r02q_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r02q_tests.log
for v in default tok_simt tokm_mb2; do
  unset MARL_MAZE_LIB
  if [ $v != default ]; then export MARL_MAZE_LIB=$PWD/variants/$v.so; fi
  timeout 300 python tools/tok_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r02q_tok.jsonl
done
unset MARL_MAZE_LIB
timeout 300 python tools/k4_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r02q_k4.jsonl
