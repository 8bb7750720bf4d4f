#!/bin/bash
# ad-hoc GPU probe used during development
mkdir -p gpurun_out/r02a
( timeout 600 python tests/gpu_tools/warp_debug.py small ) > gpurun_out/r02a/warp_debug.log 2>&1
tail -12 gpurun_out/r02a/warp_debug.log
( timeout 300 python tools/variant_bench.py own kernel=2 ) > gpurun_out/r02a/variant_k2.log 2>&1; tail -2 gpurun_out/r02a/variant_k2.log
( timeout 300 python tools/variant_bench.py own kernel=1 ) > gpurun_out/r02a/variant_k1.log 2>&1; tail -1 gpurun_out/r02a/variant_k1.log
( GCKPP_B200_LIB=geos_chem_b200/libgckpp_b200_prof.so GCKPP_PROFILE=1 timeout 120 python tools/smem_one.py 444 2 ) > gpurun_out/r02a/prof.log 2>&1; tail -3 gpurun_out/r02a/prof.log
( timeout 600 python tests/gpu_tools/gpu_check.py 20000 warm kernel=2 ) > gpurun_out/r02a/gpu_check.log 2>&1; tail -12 gpurun_out/r02a/gpu_check.log
( timeout 300 compute-sanitizer --tool memcheck python tools/smem_one.py 12 2 ) > gpurun_out/r02a/memcheck.log 2>&1; tail -4 gpurun_out/r02a/memcheck.log
