#!/bin/bash
# ncu capture of the tcgen05 policy GEMM.  Usage: bash tools/gpu_ncu_k4.sh tag
TAG=${1:-x}
mkdir -p gpurun_out
CMD="python tools/rollout_bench.py --envs 65536 --horizon 4 --epochs 1 --no-update"
timeout 300 $CMD > gpurun_out/plain_k4_$TAG.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_linear_tf32x3 -s 3 -c 3 -f -o gpurun_out/k4_$TAG $CMD > gpurun_out/ncu_k4_$TAG.log 2>&1
echo "ncu rc=$?"; tail -1 gpurun_out/plain_k4_$TAG.log | cut -c1-400
