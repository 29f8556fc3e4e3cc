#!/bin/bash
# r02f: state recovery after the container was re-created: full GPU tests, GEMM kernel timings, K4 forward (3xFP16 vs 3xTF32), rollout, launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02f_tests.log 2>&1; echo "all tests rc=$?"; tail -8 gpurun_out/r02f_tests.log
timeout 300 python tools/k5_bench.py --skip-update 2>&1 | tail -1 | tee gpurun_out/r02f_k5_kernels.json
for v in default tf32; do
  unset MARL_MAZE_TF32_TRUNK
  if [ $v = tf32 ]; then export MARL_MAZE_TF32_TRUNK=1; fi
  echo "== $v" | tee -a gpurun_out/r02f_k4.jsonl
  timeout 300 python tools/k4_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r02f_k4.jsonl
done
unset MARL_MAZE_TF32_TRUNK
timeout 600 python tools/rollout_bench.py --epochs 3 2>&1 | tail -1 | tee gpurun_out/r02f_rollout.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02f_k4_launches.csv python tools/k4_bench.py > gpurun_out/r02f_k4_ncu.log 2>&1
echo "launch list rc=$?"
