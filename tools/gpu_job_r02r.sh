#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tokens_mma" -s 3 -c 1 -f -o gpurun_out/tokm_r02r python tools/tok_bench.py > gpurun_out/r02r_ncu.log 2>&1
echo "ncu rc=$?"
