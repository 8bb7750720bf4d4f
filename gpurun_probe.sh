#!/bin/bash
o=gpurun_out/r02s; mkdir -p $o
( GCKPP_B200_LIB=geos_chem_b200/libgckpp_b200_prof.so GCKPP_PROFILE=1 timeout 300 python tools/smem_one.py 9472 3 ) > $o/lane_prof2.log 2>&1; tail -2 $o/lane_prof2.log
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $o/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $o/pytest_gpu.log; tail -5 $o/pytest_gpu.log
( VB_ITERS=1 timeout 300 python tools/variant_bench.py own kernel=3 ) > $o/variant_k3.log 2>&1; tail -1 $o/variant_k3.log
