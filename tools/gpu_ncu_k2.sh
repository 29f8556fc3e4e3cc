#!/bin/bash
# ncu full capture of K2 only (after the same command ran clean). Usage: bash tools/gpu_ncu_k2.sh tag
TAG=${1:-x}
mkdir -p gpurun_out
CMD="python tools/k2_bench.py --steps 6 --warmup 6"
timeout 300 $CMD > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_step_obs -s 8 -c 2 -f -o gpurun_out/k2_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/plain_$TAG.log
