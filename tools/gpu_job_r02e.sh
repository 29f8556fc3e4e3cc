#!/bin/bash
# r02e: where the time goes in the 3xFP16 GEMM and the R-row token kernel: kernel timings, ncu launch list, ncu --set full
mkdir -p gpurun_out
timeout 300 python tools/k5_bench.py --skip-update 2>&1 | tail -1 | tee gpurun_out/r02e_k5_kernels.json
timeout 300 python tools/k4_bench.py > gpurun_out/r02e_k4_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02e_k4_launches.csv python tools/k4_bench.py > gpurun_out/r02e_k4_ncu.log 2>&1
echo "launch list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_linear_f16x3" -s 6 -c 3 -f -o gpurun_out/k4f16_r02e python tools/k4_bench.py > gpurun_out/r02e_ncu_f16.log 2>&1
echo "ncu f16 rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_tokens" -s 2 -c 1 -f -o gpurun_out/tok_r02e python tools/k4_bench.py > gpurun_out/r02e_ncu_tok.log 2>&1
echo "ncu tok rc=$?"
