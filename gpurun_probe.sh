#!/bin/bash
o=gpurun_out/r02y; mkdir -p $o
for v in _u4 _u6 _u8; do ( GCKPP_B200_LIB=geos_chem_b200/libgckpp_b200$v.so timeout 300 python bench.py --config 5-hg --steps 5 --warmup 3 --no-cpu-baseline ) 2>&1 | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('variant $v', d['value'], d['e2e']['value'], d['roofline']['kernel_ms'])" | tee -a $o/hg_variants.log; done
