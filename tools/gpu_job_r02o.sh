#!/bin/bash
# r02o: launch list of the policy forward with the fused trunk (uniform-register MMA issue, cluster pair)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max --clock-control none -k regex:"k_tokens|k_linear|k_heads|k_critic|k_trunk" -s 12 -c 6 --csv --log-file gpurun_out/r02o_k4_launches.csv python tools/k4_bench.py > gpurun_out/r02o_k4_ncu.log 2>&1
echo "launch list rc=$?"
