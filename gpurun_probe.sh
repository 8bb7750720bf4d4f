#!/bin/bash
o=gpurun_out/r02ak; mkdir -p $o
for v in lmax7 lmax11 lmax4 solve3 solve11; do
( GCKPP_B200_LIB=$PWD/geos_chem_b200/libgckpp_b200_$v.so VB_ITERS=3 timeout 300 python tools/variant_bench.py own ) > $o/variant_$v.log 2>&1; echo $v; tail -n 1 $o/variant_$v.log
done
( VB_ITERS=3 timeout 300 python tools/variant_bench.py own ) > $o/variant_default.log 2>&1; tail -n 1 $o/variant_default.log
