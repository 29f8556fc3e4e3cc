#!/bin/bash
# r02d: first run of the 3xFP16 two-CTA/SM GEMM (mm_linear16.cu) and the R-rows-per-warp token kernel
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_update_gpu.py -x -q -k "f16x3" > gpurun_out/r02d_gemm_test.log 2>&1; echo "gemm test rc=$?"; tail -30 gpurun_out/r02d_gemm_test.log
timeout 600 python -m pytest tests/test_policy_gpu.py -q > gpurun_out/r02d_policy_test.log 2>&1; echo "policy tests rc=$?"; tail -30 gpurun_out/r02d_policy_test.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02d_tests.log 2>&1; echo "all tests rc=$?"; tail -15 gpurun_out/r02d_tests.log
for v in default tf32 tok_r1 tok_r2_mb3 tok_r3_mb3 tok_r4_mb3 tok_r4_mb2; do
  unset MARL_MAZE_LIB MARL_MAZE_TF32_TRUNK
  if [ $v = tf32 ]; then export MARL_MAZE_TF32_TRUNK=1; elif [ $v != default ]; then export MARL_MAZE_LIB=$PWD/variants/$v.so; fi
  echo "== $v" | tee -a gpurun_out/r02d_k4.jsonl
  timeout 300 python tools/k4_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r02d_k4.jsonl
done
unset MARL_MAZE_LIB MARL_MAZE_TF32_TRUNK
timeout 600 python tools/rollout_bench.py --epochs 3 2>&1 | tail -1 | tee gpurun_out/r02d_rollout.json
