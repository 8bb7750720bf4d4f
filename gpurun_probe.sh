#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "pipelined or fixture_replicated or active_mask" 2>&1 | tail -4
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',d['value'],'e2e',d['e2e'])"
