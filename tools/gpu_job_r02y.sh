#!/bin/bash
# r02y (2 GPUs): NCCL gradient test; 300 iterations at the reference settings with 4096 mazes per rank (the update's CUDA graphs contain the NCCL all-reduces)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_distributed_gpu.py -x -q > gpurun_out/r02y_dist_test.log 2>&1; echo "dist test rc=$?"; tail -4 gpurun_out/r02y_dist_test.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/train_demo.py --envs 4096 --horizon 128 --iters 300 --side-half 13 --max-t 1200 --lr 0.00014 2>&1 | tail -3 | tee gpurun_out/r02y_train_300_n2.txt
