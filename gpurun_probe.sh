#!/bin/bash
mkdir -p gpurun_out/r01c
( timeout 300 python tools/smem_debug.py small ) 2>&1 | tee gpurun_out/r01c/small.log | tail -3
( timeout 300 python tools/smem_debug.py full ) 2>&1 | tee gpurun_out/r01c/full.log | tail -2
timeout 120 python tools/hg_debug.py 2>&1 | tail -3
