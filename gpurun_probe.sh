#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "other_rosenbrock" -s 2>&1 | tail -12
