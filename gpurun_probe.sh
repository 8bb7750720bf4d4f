#!/bin/bash
o=gpurun_out/r02ap; mkdir -p $o
for v in p0 p1 p3; do
( GCKPP_B200_LIB=$PWD/geos_chem_b200/libgckpp_b200_$v.so GCKPP_PROFILE=1 timeout 300 python tools/smem_one.py 444 1 ) > $o/prof_$v.log 2>&1; echo $v; cat $o/prof_$v.log | cut -c1-900
done
