#!/bin/bash
o=gpurun_out/r02d; mkdir -p $o
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ros_warp -c 1 -o $o/ros_warp_full -f \
    python tools/variant_bench.py own kernel=2 > $o/ncu_full.log 2>&1
ncu -i $o/ros_warp_full.ncu-rep --page raw --csv > $o/ros_warp_raw.csv 2>/dev/null
ncu -i $o/ros_warp_full.ncu-rep --page details --csv > $o/ros_warp_details.csv 2>/dev/null
ncu -i $o/ros_warp_full.ncu-rep --page source --csv > $o/ros_warp_source.csv 2>/dev/null
ls -la $o; tail -3 $o/ncu_full.log
