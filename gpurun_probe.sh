#!/bin/bash
# bring-up of the shared-memory kernel: small -> grid -> full, each under its own timeout
mkdir -p gpurun_out/r01b
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
( timeout 300 python tools/smem_debug.py small ) 2>&1 | tee gpurun_out/r01b/small.log
( timeout 300 python tools/smem_debug.py full ) 2>&1 | tee gpurun_out/r01b/full.log
