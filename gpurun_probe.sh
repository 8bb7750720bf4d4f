#!/bin/bash
o=gpurun_out/r02e; mkdir -p $o
( timeout 600 python tests/gpu_tools/warp_debug.py grid ) > $o/warp_debug.log 2>&1
tail -6 $o/warp_debug.log
( timeout 300 python tools/variant_bench.py own kernel=2 ) > $o/variant.log 2>&1; echo "variant: $(tail -1 $o/variant.log)"
( GCKPP_B200_LIB=geos_chem_b200/libgckpp_b200_prof.so GCKPP_PROFILE=1 timeout 120 python tools/smem_one.py 444 2 ) > $o/prof.log 2>&1; tail -6 $o/prof.log
