#!/bin/bash
# r02k: ncu --set full of the fused trunk kernel
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_trunk_fused" -s 3 -c 1 -f -o gpurun_out/trunk_r02k python tools/k4_bench.py > gpurun_out/r02k_ncu.log 2>&1
echo "ncu rc=$?"
