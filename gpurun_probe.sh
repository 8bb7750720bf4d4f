#!/bin/bash
# ad-hoc GPU probe used during development: parity of the default kernel on a small sample + one timing line
( timeout 300 python tests/gpu_tools/smem_debug.py small ) 2>&1 | tail -3
timeout 200 python tools/variant_bench.py own 2>&1 | tail -1
