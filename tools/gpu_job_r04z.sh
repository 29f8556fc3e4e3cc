#!/bin/bash
# critic once per rollout over the whole observation buffer (default) against the per-step side-stream launch
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ppo_gpu.py -q 2>&1 | tail -3
for bv in 1 0 1 0; do
  MARL_MAZE_BATCHED_VALUES=$bv timeout 600 python tools/rollout_bench.py --epochs 4 --no-update 2>&1 | tail -1 | python -c "
import json,sys,os
d=json.loads(sys.stdin.read()); print(json.dumps({'batched_values': os.environ.get('MARL_MAZE_BATCHED_VALUES'), 'rollout_ms': [round(e['rollout_ms'],2) for e in d['epochs']]}))" | tee -a gpurun_out/r04z_rollout.jsonl
done
