#!/bin/bash
o=gpurun_out/r02t; mkdir -p $o
( time timeout 1500 python -m pytest tests -m gpu -x -q -k "retry or option_errors or two_devices or pipelined" ) > $o/pytest_new.log 2>&1; tail -6 $o/pytest_new.log
( time timeout 900 python bench.py --steps 3 --warmup 3 ) > $o/bench.log 2>&1; tail -1 $o/bench.log | cut -c1-3000
( time timeout 900 python bench.py --impl reference --steps 1 --warmup 0 ) > $o/bench_ref.log 2>&1; tail -1 $o/bench_ref.log | cut -c1-600
