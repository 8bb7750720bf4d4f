#!/bin/bash
# ad-hoc GPU probe used during development
( timeout 300 python tests/gpu_tools/smem_debug.py small ) 2>&1 | tail -3
timeout 200 python tools/variant_bench.py own 2>&1 | tail -1
timeout 120 python tests/gpu_tools/hg_debug.py 2>&1 | tail -2
