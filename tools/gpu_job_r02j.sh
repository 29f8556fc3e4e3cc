#!/bin/bash
# r02j: first run of the fused trunk kernel (mm_trunk_fused.cu): parity tests, then timing against the per-layer kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_policy_gpu.py -x -q -k "fused_trunk" > gpurun_out/r02j_fused_test.log 2>&1; echo "fused tests rc=$?"; tail -30 gpurun_out/r02j_fused_test.log
timeout 300 python tools/k4_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r02j_k4.jsonl
timeout 300 python tools/k4_bench.py --no-fused 2>&1 | tail -1 | tee -a gpurun_out/r02j_k4.jsonl
timeout 300 python tools/k4_bench.py --envs 4096 2>&1 | tail -1 | tee -a gpurun_out/r02j_k4.jsonl
timeout 300 python tools/k4_bench.py --envs 4096 --no-fused 2>&1 | tail -1 | tee -a gpurun_out/r02j_k4.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"k_tokens|k_linear|k_heads|k_critic|k_trunk" -s 12 -c 6 --csv --log-file gpurun_out/r02j_k4_launches.csv python tools/k4_bench.py > gpurun_out/r02j_k4_ncu.log 2>&1
echo "launch list rc=$?"
