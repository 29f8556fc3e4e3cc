#!/bin/bash
# r02h: ncu --set full of the 3xFP16 GEMM (layer 0 and layer 1) and the token kernel at 131072 rows
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_linear_f16x3" -s 6 -c 2 -f -o gpurun_out/k4f16_r02h python tools/k4_bench.py > gpurun_out/r02h_ncu_f16.log 2>&1
echo "ncu f16 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tokens" -s 2 -c 1 -f -o gpurun_out/tok_r02h python tools/k4_bench.py > gpurun_out/r02h_ncu_tok.log 2>&1
echo "ncu tok rc=$?"
