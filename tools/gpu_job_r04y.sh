#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_generator_gpu.py tests/test_policy_gpu.py -q 2>&1 | tail -6
