#!/bin/bash
# r02c: full GPU test suite (new illegal-action / KAT-5 tests) + both bench arms as the driver runs them
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02c_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r02c_tests.log
tail -15 gpurun_out/r02c_tests.log
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r02c_bench_ref.json 2> gpurun_out/r02c_bench_ref.err; echo "ref rc=$?"
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/r02c_bench.err
