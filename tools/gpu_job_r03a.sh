#!/bin/bash
# r03a: rectangular mazes, agent setters, everything else: full GPU suite
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r03a_tests.log 2>&1; echo "all tests rc=$?"; tail -12 gpurun_out/r03a_tests.log
