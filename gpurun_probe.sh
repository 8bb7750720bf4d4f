#!/bin/bash
o=gpurun_out/r02af; mkdir -p $o
( VB_ITERS=3 timeout 300 python tools/variant_bench.py own ) > $o/variant_bank.log 2>&1; tail -2 $o/variant_bank.log
( GCKPP_B200_LIB=$PWD/geos_chem_b200/libgckpp_b200_nobank.so VB_ITERS=3 timeout 300 python tools/variant_bench.py own ) > $o/variant_nobank.log 2>&1; tail -2 $o/variant_nobank.log
( time timeout 1200 python -m pytest tests -m gpu -x -q -s ) > $o/pytest_gpu.log 2>&1; grep -E "auto-reduce|device het|passed|failed|Error|^E |different steps" $o/pytest_gpu.log | head -30
