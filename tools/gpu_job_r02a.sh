#!/bin/bash
# r02a: K2 window transport A/B (bulk vs per-row cp.async, occupancy variants) + full GPU test suite on the default library
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02a_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r02a_tests.log
tail -3 gpurun_out/r02a_tests.log
for v in default k2_ldgsts k2_bulk_mb5 k2_ldgsts_mb5 k2_bulk_mb6; do
  if [ $v = default ]; then unset MARL_MAZE_LIB; else export MARL_MAZE_LIB=$PWD/variants/$v.so; fi
  timeout 300 python tools/k2_bench.py --steps 200 --warmup 40 2>&1 | tail -1 | tee -a gpurun_out/r02a_k2.jsonl
done
unset MARL_MAZE_LIB
timeout 300 python tools/k2_bench.py --side-half 13 --steps 200 --warmup 40 2>&1 | tail -1 | tee -a gpurun_out/r02a_k2.jsonl
timeout 600 python tools/rollout_bench.py --no-update --epochs 3 2>&1 | tail -1 | tee gpurun_out/r02a_rollout.json
