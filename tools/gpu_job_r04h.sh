#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/k1_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r04h_k1.jsonl
timeout 300 python tools/k1_bench.py --mazes 393216 --side-half 13 2>&1 | tail -1 | tee -a gpurun_out/r04h_k1.jsonl
timeout 300 python tools/k1_bench.py --mazes 262144 --side-half 13 --difficulty 4 2>&1 | tail -1 | tee -a gpurun_out/r04h_k1.jsonl
