#!/bin/bash
# r02g: FFMA2 microbenchmark + launch list of the policy forward (our kernels only)
mkdir -p gpurun_out
./tools/micro/ffma2 | tee gpurun_out/r02g_ffma2.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_tokens|k_linear|k_heads|k_critic" -c 60 --csv --log-file gpurun_out/r02g_k4_launches.csv python tools/k4_bench.py > gpurun_out/r02g_k4_ncu.log 2>&1
echo "launch list rc=$?"
