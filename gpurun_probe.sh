#!/bin/bash
bash tools/gpu_round.sh r02as
o=gpurun_out/r02as
( timeout 600 python bench.py --config 1 --steps 3 --warmup 3 --no-cpu-baseline ) > $o/bench_config1.log 2>&1; tail -n 1 $o/bench_config1.log | cut -c1-200
( timeout 600 python bench.py --hstart cold --steps 2 --warmup 3 --no-cpu-baseline ) > $o/bench_cold.log 2>&1; tail -n 1 $o/bench_cold.log | cut -c1-200
( timeout 900 python bench.py --config 5-ar --steps 3 --warmup 3 ) > $o/bench_ar.log 2>&1; tail -n 1 $o/bench_ar.log | cut -c1-200
