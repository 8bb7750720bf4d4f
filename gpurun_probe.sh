#!/bin/bash
o=gpurun_out/r02o; mkdir -p $o
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $o/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $o/pytest_gpu.log
tail -5 $o/pytest_gpu.log
( time timeout 600 python bench.py --steps 3 --warmup 3 ) > $o/bench.log 2>&1; tail -2 $o/bench.log
