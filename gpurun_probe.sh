#!/bin/bash
( timeout 300 python tests/gpu_tools/smem_debug.py small ) 2>&1 | tail -3
timeout 200 python tools/variant_bench.py own 2>&1 | tail -1
GCKPP_B200_LIB=$PWD/geos_chem_b200/libvar_nw16.so timeout 300 python tests/gpu_tools/smem_debug.py small 2>&1 | tail -3
GCKPP_B200_LIB=$PWD/geos_chem_b200/libvar_nw16.so timeout 200 python tools/variant_bench.py own 2>&1 | tail -1
