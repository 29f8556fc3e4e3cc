#!/bin/bash
# single-env loop with the logits prefetched by the step (one synchronisation per env step)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ppo_gpu.py -x -q 2>&1 | tail -5
for i in 1 2; do timeout 300 python tools/single_env_bench.py 2>&1 | tail -1 | tee -a gpurun_out/r04d_single_env.jsonl; done
timeout 300 python examples/main.py 2>&1 | tail -5
