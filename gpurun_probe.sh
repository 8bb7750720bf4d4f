#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "option_variants" -s 2>&1 | tail -14
