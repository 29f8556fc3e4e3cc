#!/bin/bash
# mm_gather_rows in the update: tests, config-3 iteration
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_update_gpu.py tests/test_ppo_gpu.py -q 2>&1 | tail -3
timeout 600 python tools/rollout_bench.py --epochs 4 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(json.dumps({'rollout_ms': [round(e['rollout_ms'],2) for e in d['epochs']], 'update_ms': [round(e['update_ms'],1) for e in d['epochs']]}))" | tee gpurun_out/r05f_rollout_update.json
