#!/bin/bash
o=gpurun_out/r02aa; mkdir -p $o
( time timeout 900 python -m pytest tests -m gpu -x -q -s -k "heterogeneous or update_rconst or fixture_replicated" ) > $o/pytest_het.log 2>&1; grep -E "device het|passed|failed|Error|^E " $o/pytest_het.log | head -20
