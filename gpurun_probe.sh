#!/bin/bash
o=gpurun_out/r02aq; mkdir -p $o
( VB_ITERS=3 timeout 300 python tools/variant_bench.py own ) > $o/variant_recip.log 2>&1; tail -n 1 $o/variant_recip.log
( GCKPP_B200_LIB=$PWD/geos_chem_b200/libgckpp_b200_norecip.so VB_ITERS=3 timeout 300 python tools/variant_bench.py own ) > $o/variant_norecip.log 2>&1; tail -n 1 $o/variant_norecip.log
( time timeout 1200 python -m pytest tests -m gpu -x -q -s ) > $o/pytest_gpu.log 2>&1; grep -E "passed|failed|Error|^E |different steps [1-9]" $o/pytest_gpu.log | head -20
( timeout 900 python bench.py --steps 3 --warmup 3 ) > $o/bench.log 2>&1; grep '^{' $o/bench.log | tail -n 1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], {k:v for k,v in d['parity'].items() if 'hist' not in k})"
