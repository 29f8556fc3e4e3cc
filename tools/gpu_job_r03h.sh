#!/bin/bash
# r03h: ncu --set full of K2 in RESET STEADY STATE (inside bench.py's timed region, after the phase spreader), for the DRAM bytes per launch
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step_obs -s 2450 -c 2 -f -o gpurun_out/k2_steady_r03h python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-legs > gpurun_out/r03h_ncu.log 2>&1; echo "ncu rc=$?"
