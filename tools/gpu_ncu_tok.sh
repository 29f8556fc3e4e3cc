#!/bin/bash
# ncu capture of the token kernel.  Usage: bash tools/gpu_ncu_tok.sh tag [variant.so]
TAG=${1:-x}
[ -n "$2" ] && export MARL_MAZE_LIB=$PWD/variants/$2
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tokens -s 3 -c 1 -f -o gpurun_out/tok_$TAG python tools/tok_bench.py > gpurun_out/ncu_tok_$TAG.log 2>&1
echo "ncu rc=$?"
