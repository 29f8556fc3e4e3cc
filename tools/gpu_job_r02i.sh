#!/bin/bash
# r02i: token kernel variants after the tile-stride fix
mkdir -p gpurun_out
for v in default tok_r1 tok_r2_ts32 tok_r2_mb3 tok_r3_mb3 tok_r4_mb2; do
  unset MARL_MAZE_LIB
  if [ $v != default ]; then export MARL_MAZE_LIB=$PWD/variants/$v.so; fi
  timeout 300 python tools/tok_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r02i_tok.jsonl
done
