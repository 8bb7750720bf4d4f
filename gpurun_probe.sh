#!/bin/bash
o=gpurun_out/r02ac; mkdir -p $o
( time timeout 1200 python -m pytest tests -m gpu -x -q -s -k "autoreduce or heterogeneous" ) > $o/pytest_ar_het.log 2>&1; grep -E "auto-reduce|device het|passed|failed|Error|^E " $o/pytest_ar_het.log | head -30
( time timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline ) > $o/bench.log 2>&1; tail -4 $o/bench.log | cut -c1-300
( time timeout 600 python bench.py --config 5-ar --steps 2 --warmup 1 --no-cpu-baseline ) > $o/bench_ar.log 2>&1; tail -4 $o/bench_ar.log | cut -c1-900
