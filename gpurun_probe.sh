#!/bin/bash
( timeout 300 python tools/smem_debug.py small ) 2>&1 | tail -3
timeout 200 python tools/variant_bench.py own 2>&1 | tail -1
