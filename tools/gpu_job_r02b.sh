#!/bin/bash
# r02b: K2 (bulk windows, 5 blocks/SM) occupancy / pipelining variants, then ncu --set full of the default K2
mkdir -p gpurun_out
for v in default k2_g2_mb3 k2_g2_mb4 k2_t64_mb10 k2_t64_mb11 k2_t96_mb7; do
  if [ $v = default ]; then unset MARL_MAZE_LIB; else export MARL_MAZE_LIB=$PWD/variants/$v.so; fi
  timeout 300 python tools/k2_bench.py --steps 200 --warmup 40 2>&1 | tail -1 | tee -a gpurun_out/r02b_k2.jsonl
done
unset MARL_MAZE_LIB
timeout 300 python tools/k2_bench.py --side-half 13 --steps 200 --warmup 40 2>&1 | tail -1 | tee -a gpurun_out/r02b_k2.jsonl
timeout 300 python tools/k2_bench.py --envs 65536 --side-half 13 --steps 400 --warmup 40 2>&1 | tail -1 | tee -a gpurun_out/r02b_k2.jsonl
bash tools/gpu_ncu_k2.sh r02b
