#!/bin/bash
# r04x (2 GPUs): the NCCL gradient test, the bench at N = 2 (train_iter + strong legs)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_distributed_gpu.py -x -q > gpurun_out/r04x_dist_test.log 2>&1; echo "dist test rc=$?"; tail -12 gpurun_out/r04x_dist_test.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r04x_bench_n2.json 2> gpurun_out/r04x_bench_n2.err; echo "bench rc=$?"; tail -3 gpurun_out/r04x_bench_n2.err; cut -c1-3000 gpurun_out/r04x_bench_n2.json
