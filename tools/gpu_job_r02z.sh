#!/bin/bash
# r02z (2 GPUs): 300 iterations at the reference settings, 4096 mazes per rank, update graphs segmented around the NCCL calls.  Short timeouts.
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/train_demo.py --envs 4096 --horizon 128 --iters 300 --side-half 13 --max-t 1200 --lr 0.00014 2>&1 | tail -3 | tee gpurun_out/r02z_train_300_n2.txt
echo "train rc=$?"
