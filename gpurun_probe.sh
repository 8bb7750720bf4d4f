#!/bin/bash
o=gpurun_out/r02u; mkdir -p $o
( time timeout 900 python -m pytest tests -m gpu -x -q -k "retry or option_errors or two_devices" ) > $o/pytest_new.log 2>&1; tail -4 $o/pytest_new.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 ) > $o/bench_n2.log 2>&1; grep "^{" $o/bench_n2.log | cut -c1-700
