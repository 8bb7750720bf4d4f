#!/bin/bash
# last visit of round 2: the driver's smoke() on the final build
o=gpurun_out/r02ba; mkdir -p $o
( time timeout 100 python -c "import __graft_entry__ as g; g.smoke()" ) > $o/smoke.log 2>&1; tail -n 5 $o/smoke.log
