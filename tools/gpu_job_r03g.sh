#!/bin/bash
# r03g: K2 with the dir-to-exit direction derived from the move (field read only when moving along the route): parity, timing early / steady state
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_env_parity_gpu.py tests/test_generator_gpu.py -x -q > gpurun_out/r03g_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r03g_tests.log
timeout 300 python tools/k2_bench.py 2>&1 | tail -1 | tee gpurun_out/r03g_k2.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-legs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps({k:d[k] for k in ('value','ms_per_step')}), d['roofline']['frac'], d['e2e']['value'])" | tee gpurun_out/r03g_bench_short.txt
