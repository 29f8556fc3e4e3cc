#!/bin/bash
# r03b: Adam variants against fp64; final ncu captures (fused trunk, token kernel, K2), launch list of the bench command
mkdir -p gpurun_out
timeout 120 python tools/adam_divergence.py | tee gpurun_out/r03b_adam_divergence.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_trunk_fused" -s 3 -c 1 -f -o gpurun_out/trunk_r03b python tools/k4_bench.py > gpurun_out/r03b_ncu_trunk.log 2>&1; echo "ncu trunk rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tokens_mma" -s 3 -c 1 -f -o gpurun_out/tokm_r03b python tools/k4_bench.py > gpurun_out/r03b_ncu_tok.log 2>&1; echo "ncu tok rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_step_obs" -s 40 -c 2 -f -o gpurun_out/k2_r03b python tools/k2_bench.py > gpurun_out/r03b_ncu_k2.log 2>&1; echo "ncu k2 rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r03b_bench_plain.json 2> gpurun_out/r03b_bench_plain.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" -c 400 --csv --log-file gpurun_out/r03b_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-legs > gpurun_out/r03b_ncu_bench.log 2>&1; echo "launch list rc=$?"
