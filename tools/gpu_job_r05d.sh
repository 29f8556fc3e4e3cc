#!/bin/bash
# K2: L2 prefetch of the state lines of the warp N groups ahead (about one block lifetime)
mkdir -p gpurun_out
for v in default k2_pf1480 k2_pf2960 k2_pf5920 default k2_pf1480 k2_pf2960 k2_pf5920; do
  unset MARL_MAZE_LIB
  if [ $v != default ]; then export MARL_MAZE_LIB=$PWD/variants/$v.so; fi
  timeout 300 python tools/k2_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r05d_k2.jsonl
done
