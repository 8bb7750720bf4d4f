#!/bin/bash
# ad-hoc GPU probe used during development
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "carbon" 2>&1 | tail -8
