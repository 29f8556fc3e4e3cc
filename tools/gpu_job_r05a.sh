#!/bin/bash
# config 5 on the current code: 300 PPO iterations at the reference settings, 4096 mazes, one GPU
mkdir -p gpurun_out
timeout 900 python tools/train_demo.py --envs 4096 --horizon 128 --iters 300 --side-half 13 --max-t 1200 --lr 0.00014 2>&1 | tail -4 | tee gpurun_out/r05a_train_300.txt
