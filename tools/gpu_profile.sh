#!/bin/bash
# Run on the GPU box (via gpurun): tests, smoke, bench, then the ncu launch list and one full capture of K2.
# Usage: bash tools/gpu_profile.sh [tag]      outputs land in gpurun_out/
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
SHORT="python bench.py --steps 8 --warmup 3 --e2e-steps 8 --no-cpu-baseline"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 600 python bench.py --impl reference --steps 64 --warmup 4 > gpurun_out/bench_ref_$TAG.json 2>&1; echo "ref rc=$?"; cat gpurun_out/bench_ref_$TAG.json
timeout 300 $SHORT > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv $SHORT > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
timeout 300 $SHORT > gpurun_out/plain2_$TAG.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_step_obs -s 6 -c 2 -f -o gpurun_out/k2_$TAG $SHORT > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out | tail -20
