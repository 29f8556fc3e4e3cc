#!/bin/bash
# r05k (8 GPUs): the bench at N = 8 as the driver launches it, on the final code.  Tight timeout.
mkdir -p gpurun_out
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 ) > gpurun_out/r05k_bench_n8.json 2> gpurun_out/r05k_bench_n8.err; echo "bench rc=$?"; tail -4 gpurun_out/r05k_bench_n8.err
