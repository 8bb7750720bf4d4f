#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list, one full ncu capture of the integrator.
# Usage (from the repo root on the box): bash tools/gpu_round.sh <tag> [bench options]
tag=${1:-r01}; shift
out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
nproc > $out/nproc.txt; lscpu | head -20 >> $out/nproc.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $out/pytest_gpu.log
( time timeout 900 python bench.py --steps 3 --warmup 3 "$@" ) > $out/bench.log 2>&1
( time timeout 600 python bench.py --impl reference --steps 1 --warmup 0 --ref-cells 20000 ) > $out/bench_ref.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline "$@" > $out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ros_ -c 1 -o $out/ros_full -f \
    python bench.py --steps 1 --warmup 0 --cells 47360 --no-cpu-baseline "$@" > $out/ncu_full.log 2>&1
ncu -i $out/ros_full.ncu-rep --page raw --csv > $out/ros_full_raw.csv 2>/dev/null
ncu -i $out/ros_full.ncu-rep --page details --csv > $out/ros_full_details.csv 2>/dev/null
( timeout 600 python tests/gpu_tools/gpu_check.py 20000 warm ) > $out/gpu_check.log 2>&1
tail -3 $out/pytest_gpu.log; tail -2 $out/bench.log; tail -1 $out/bench_ref.log; tail -12 $out/gpu_check.log
