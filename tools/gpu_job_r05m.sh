#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 tools/d2h_host_path_bench.py 2>&1 | tail -1 | tee gpurun_out/r05m_d2h_n8.json
nvidia-smi topo -m 2>&1 | head -14 > gpurun_out/r05m_topo.txt; numactl -H 2>/dev/null | head -6 >> gpurun_out/r05m_topo.txt; nproc >> gpurun_out/r05m_topo.txt
