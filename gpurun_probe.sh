#!/bin/bash
o=gpurun_out/r02ab; mkdir -p $o
( time timeout 1200 python -m pytest tests -m gpu -x -q -s ) > $o/pytest_gpu.log 2>&1; grep -E "device het|passed|failed|Error|^E " $o/pytest_gpu.log | head -20
( time timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline ) > $o/bench.log 2>&1; tail -1 $o/bench.log | cut -c1-600
