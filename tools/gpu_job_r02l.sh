#!/bin/bash
# r02l: fused trunk v2 (hi plane in TMEM, 8-stage weight ring, heads as a fourth layer): parity, role profile, timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_policy_gpu.py -x -q > gpurun_out/r02l_policy_test.log 2>&1; echo "policy tests rc=$?"; tail -15 gpurun_out/r02l_policy_test.log
MARL_MAZE_LIB=$PWD/variants/tf_prof.so timeout 300 python tools/trunk_profile.py | tee -a gpurun_out/r02l_trunk_profile.json
timeout 300 python tools/k4_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r02l_k4.jsonl
timeout 300 python tools/k4_bench.py --envs 4096 2>&1 | tail -1 | tee -a gpurun_out/r02l_k4.jsonl
