#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "autoreduce" -s 2>&1 | tail -12
timeout 200 python tools/variant_bench.py own 2>&1 | tail -1
