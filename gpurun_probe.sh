#!/bin/bash
tag=r02ay; out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
nproc > $out/nproc.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $out/pytest_gpu.log
( time timeout 900 python bench.py --steps 3 --warmup 3 ) > $out/bench.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ros_ -c 1 -o $out/ros_full -f python bench.py --steps 1 --warmup 0 --cells 47360 --no-cpu-baseline > $out/ncu_full.log 2>&1
ncu -i $out/ros_full.ncu-rep --page raw --csv > $out/ros_full_raw.csv 2>/dev/null
ncu -i $out/ros_full.ncu-rep --page details --csv > $out/ros_full_details.csv 2>/dev/null
tail -3 $out/pytest_gpu.log; tail -n 1 $out/bench.log | cut -c1-300
