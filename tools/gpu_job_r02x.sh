#!/bin/bash
# r02x: full GPU suite with the graphed update; 300 iterations at the reference's settings (config 5 geometry, 4096 mazes, T = 128) on one GPU
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02x_tests.log 2>&1; echo "all tests rc=$?"; tail -6 gpurun_out/r02x_tests.log
timeout 900 python tools/train_demo.py --envs 4096 --horizon 128 --iters 300 --side-half 13 --max-t 1200 --lr 0.00014 2>&1 | tail -4 | tee gpurun_out/r02x_train_300.txt
timeout 600 python tools/rollout_bench.py --epochs 3 2>&1 | tail -1 | cut -c1-420 | tee gpurun_out/r02x_rollout.json
