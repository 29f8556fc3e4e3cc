#!/bin/bash
# r05j: what the driver runs at round end: smoke, the reference arm, our arm (timed)
mkdir -p gpurun_out
( time timeout 600 python __graft_entry__.py smoke ) > gpurun_out/r05j_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r05j_smoke.log
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r05j_bench_ref.json 2> gpurun_out/r05j_bench_ref.err; echo "ref rc=$?"; tail -3 gpurun_out/r05j_bench_ref.err
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r05j_bench.json 2> gpurun_out/r05j_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r05j_bench.err
( time timeout 1500 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/r05j_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r05j_pytest_gpu.log
