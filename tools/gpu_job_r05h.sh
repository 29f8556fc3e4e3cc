#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_env_parity_gpu.py -q -x -k "full_size" ) 2>&1 | tail -12
