#!/bin/bash
# ncu capture of the K1 generator.  Usage: bash tools/gpu_ncu_k1.sh tag
TAG=${1:-x}
mkdir -p gpurun_out
CMD="python tools/k1_bench.py --mazes 131072"
timeout 300 $CMD > gpurun_out/plain_k1_$TAG.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_generate -c 1 -f -o gpurun_out/k1_$TAG $CMD > gpurun_out/ncu_k1_$TAG.log 2>&1
echo "ncu rc=$?"; tail -1 gpurun_out/plain_k1_$TAG.log | cut -c1-400
