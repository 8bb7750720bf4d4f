#!/bin/bash
# ad-hoc GPU probe used during development: SASS-level stall samples of the production kernel
mkdir -p gpurun_out/r01j
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ros_smem -c 1 -o gpurun_out/r01j/ros_full -f \
    python bench.py --steps 1 --warmup 0 --cells 47360 --no-cpu-baseline > gpurun_out/r01j/ncu_full.log 2>&1
ncu -i gpurun_out/r01j/ros_full.ncu-rep --page source --csv > gpurun_out/r01j/ros_full_sass.csv 2>/dev/null
rm -f gpurun_out/r01j/ros_full.ncu-rep
ls -la gpurun_out/r01j
