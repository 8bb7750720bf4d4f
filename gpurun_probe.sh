#!/bin/bash
mkdir -p gpurun_out/r01e
timeout 200 python tools/variant_bench.py own 2>&1 | tail -1
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 python tools/smem_one.py 6 > gpurun_out/r01e/racecheck.log 2>&1; tail -3 gpurun_out/r01e/racecheck.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/hg_debug.py > gpurun_out/r01e/memcheck_hg.log 2>&1; tail -3 gpurun_out/r01e/memcheck_hg.log
