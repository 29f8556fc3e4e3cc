#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/d2h_host_path_bench.py 2>&1 | tail -3 | tee gpurun_out/r05l_d2h_n2.json
