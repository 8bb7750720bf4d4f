#!/bin/bash
o=gpurun_out/r02av; mkdir -p $o
( VB_ITERS=3 timeout 300 python tools/variant_bench.py own ) > $o/variant_uniform_lw.log 2>&1; tail -n 1 $o/variant_uniform_lw.log
( time timeout 1200 python -m pytest tests -m gpu -x -q -s -k "parity or fixture or autoreduce" ) > $o/pytest_gpu.log 2>&1; grep -E "passed|failed|Error|^E |different steps [1-9]" $o/pytest_gpu.log | head -20
