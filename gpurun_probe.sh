#!/bin/bash
o=gpurun_out/r02am; mkdir -p $o
( time timeout 1200 python -m pytest tests -m gpu -x -q -s -k "heterogeneous" ) > $o/pytest_het.log 2>&1; grep -E "device het|passed|failed|Error|^E " $o/pytest_het.log | head
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 3 ) > $o/bench_n2.log 2>&1; grep '^{' $o/bench_n2.log | tail -n 1 | cut -c1-700
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 ) > $o/bench_ref_n2.log 2>&1; grep '^{' $o/bench_ref_n2.log | tail -n 1 | cut -c1-500
( timeout 600 python -m pytest tests -m gpu -x -q -k "two_dev or two_gpu or devices" ) > $o/pytest_2gpu.log 2>&1; tail -n 3 $o/pytest_2gpu.log
