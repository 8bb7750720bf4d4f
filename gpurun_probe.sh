#!/bin/bash
o=gpurun_out/r02az; mkdir -p $o
( VB_ITERS=2 timeout 120 python tools/variant_bench.py own ) > $o/variant_ring_preload.log 2>&1; tail -n 1 $o/variant_ring_preload.log
( time timeout 150 python -m pytest tests -m gpu -x -q -s -k "parity or fixture or autoreduce" ) > $o/pytest_gpu.log 2>&1; grep -E "passed|failed|Error|^E |different steps [1-9]" $o/pytest_gpu.log | head -20
