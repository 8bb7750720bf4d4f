#!/bin/bash
o=gpurun_out/r02v; mkdir -p $o
( time timeout 1200 python bench.py --config 4 --shard-of 8 --steps 1 --warmup 1 --no-cpu-baseline ) > $o/bench_c180_shard.log 2>&1; tail -5 $o/bench_c180_shard.log | cut -c1-1500
( timeout 600 python bench.py --config 5-hg --steps 3 --warmup 3 ) > $o/bench_hg.log 2>&1; tail -1 $o/bench_hg.log | cut -c1-300
