#!/bin/bash
# ad-hoc GPU probe used during development
o=gpurun_out/r02c; mkdir -p $o
( timeout 600 python tests/gpu_tools/warp_debug.py small ) > $o/warp_debug.log 2>&1
tail -6 $o/warp_debug.log
for v in "" _ieee; do
  ( GCKPP_B200_LIB=geos_chem_b200/libgckpp_b200$v.so timeout 300 python tools/variant_bench.py own kernel=2 ) > $o/variant$v.log 2>&1; echo "variant '$v': $(tail -1 $o/variant$v.log)"
done
( GCKPP_B200_LIB=geos_chem_b200/libgckpp_b200_prof.so GCKPP_PROFILE=1 timeout 120 python tools/smem_one.py 444 2 ) > $o/prof.log 2>&1; tail -6 $o/prof.log
( timeout 600 python tests/gpu_tools/gpu_check.py 20000 warm kernel=2 ) > $o/gpu_check.log 2>&1; grep -E "ierr equal|kernel" $o/gpu_check.log
( timeout 400 compute-sanitizer --tool racecheck python tools/smem_one.py 12 2 ) > $o/racecheck.log 2>&1; tail -2 $o/racecheck.log
