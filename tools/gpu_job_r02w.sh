#!/bin/bash
# r02w: CUDA-graphed update steps: tests, the small-batch (config 5 geometry) iteration, the checkpoint round trip tests
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_ppo_gpu.py -x -q > gpurun_out/r02w_tests.log 2>&1; echo "ppo tests rc=$?"; tail -25 gpurun_out/r02w_tests.log
timeout 600 python tools/rollout_bench.py --envs 4096 --horizon 4 --epochs 8 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print([(round(e['rollout_ms'],2), round(e['update_ms'],2)) for e in d['epochs']])"
