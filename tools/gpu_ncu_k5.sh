#!/bin/bash
# ncu capture of the K5 update GEMMs (weight gradient + one forward-type launch).  Usage: bash tools/gpu_ncu_k5.sh tag
TAG=${1:-x}
mkdir -p gpurun_out
CMD="python tools/k5_bench.py --skip-update"
timeout 300 $CMD > gpurun_out/plain_k5_$TAG.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_wgrad_tf32x3|k_linear_tf32x3" -s 4 -c 2 -f -o gpurun_out/k5a_$TAG $CMD > gpurun_out/ncu_k5a_$TAG.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_wgrad_tf32x3 -s 2 -c 1 -f -o gpurun_out/k5_$TAG $CMD > gpurun_out/ncu_k5_$TAG.log 2>&1
echo "ncu rc=$?"; tail -1 gpurun_out/plain_k5_$TAG.log | cut -c1-300
