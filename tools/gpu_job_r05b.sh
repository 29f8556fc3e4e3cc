#!/bin/bash
# final-code profiling artefacts: ncu launch lists (bench command, one rollout), full captures of K2 (reset steady state) and of the critic
mkdir -p gpurun_out
SHORT="python bench.py --steps 8 --warmup 3 --e2e-steps 8 --no-cpu-baseline --no-extra-legs"
timeout 300 $SHORT > gpurun_out/r05b_plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r05b_bench_launches.csv $SHORT > gpurun_out/r05b_ncu_bench.log 2>&1
echo "bench launches rc=$?"
ROLL="python tools/rollout_bench.py --envs 65536 --horizon 4 --epochs 1 --no-update"
timeout 300 $ROLL > gpurun_out/r05b_plain_roll.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r05b_rollout_launches.csv $ROLL > gpurun_out/r05b_ncu_roll.log 2>&1
echo "rollout launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step_obs -s 2450 -c 2 -f -o gpurun_out/k2_steady_r05b python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-legs > gpurun_out/r05b_ncu_k2.log 2>&1; echo "k2 full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_critic -s 3 -c 1 -f -o gpurun_out/critic_r05b python tools/critic_bench.py --envs 1048576 > gpurun_out/r05b_ncu_critic.log 2>&1; echo "critic full rc=$?"
