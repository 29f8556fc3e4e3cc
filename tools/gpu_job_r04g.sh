#!/bin/bash
# K1 second generation (shared-memory planes, 2-bit move stack, stackless tree walks, compacted refill) against the first
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_generator_gpu.py tests/test_env_parity_gpu.py -x -q 2>&1 | tail -4
for v in default k1_v1; do
  unset MARL_MAZE_LIB
  if [ $v != default ]; then export MARL_MAZE_LIB=$PWD/variants/$v.so; fi
  timeout 300 python tools/k1_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r04g_k1.jsonl
  timeout 300 python tools/k1_bench.py --mazes 393216 --side-half 13 2>&1 | tail -1 | tee -a gpurun_out/r04g_k1.jsonl
  timeout 300 python tools/k1_bench.py --mazes 262144 --side-half 13 --difficulty 4 2>&1 | tail -1 | tee -a gpurun_out/r04g_k1.jsonl
done
