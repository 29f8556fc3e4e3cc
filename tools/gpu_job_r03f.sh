#!/bin/bash
# r03f (8 GPUs): the bench at N = 8 as the driver launches it, then 300 iterations at the reference settings with 4096 mazes per rank (config 5).  Tight timeouts.
mkdir -p gpurun_out
( time timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 ) > gpurun_out/r03f_bench_n8.json 2> gpurun_out/r03f_bench_n8.err; echo "bench rc=$?"; tail -4 gpurun_out/r03f_bench_n8.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 tools/train_demo.py --envs 4096 --horizon 128 --iters 300 --side-half 13 --max-t 1200 --lr 0.00014 2>&1 | tail -2 | tee gpurun_out/r03f_train_300_n8.txt
echo "train rc=$?"
