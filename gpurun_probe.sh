#!/bin/bash
o=gpurun_out/r02ax; mkdir -p $o
for v in e4 e7 e15; do
( GCKPP_B200_LIB=$PWD/geos_chem_b200/libgckpp_b200_$v.so VB_ITERS=2 timeout 300 python tools/variant_bench.py own ) > $o/variant_$v.log 2>&1; echo $v; tail -n 1 $o/variant_$v.log
done
( VB_ITERS=2 timeout 300 python tools/variant_bench.py own ) > $o/variant_default.log 2>&1; tail -n 1 $o/variant_default.log
