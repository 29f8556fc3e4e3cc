#!/bin/bash
# full GPU suite on the third-generation token kernel (rollout) + maps-only forward for the update; token kernel timings both ways; ncu of the new kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu 2>&1 | tail -4
timeout 300 python tools/tok_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r04w_tok.jsonl
timeout 300 python tools/tok_bench.py --maps-only 2>&1 | tail -1 | tee -a gpurun_out/r04w_tok.jsonl
timeout 600 python tools/rollout_bench.py --epochs 3 2>&1 | tail -1 | tee gpurun_out/r04w_rollout.json
bash tools/gpu_ncu_tok.sh r04w
