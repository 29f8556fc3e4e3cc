#!/bin/bash
mkdir -p gpurun_out
ROLL="python tools/rollout_bench.py --envs 65536 --horizon 4 --epochs 1 --no-update"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_tokens|k_trunk|k_critic|k_step_obs|k_gae|k_generate|k_add" -c 60 --csv --log-file gpurun_out/r05c_rollout_launches.csv $ROLL > gpurun_out/r05c_ncu_roll.log 2>&1
echo "rollout launches rc=$?"
