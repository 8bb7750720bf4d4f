#!/bin/bash
o=gpurun_out/r02w; mkdir -p $o
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --config 4 --steps 1 --warmup 1 ) > $o/bench_c180_n8.log 2>&1; grep "^{" $o/bench_c180_n8.log | cut -c1-900; tail -3 $o/bench_c180_n8.log | cut -c1-200
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 2 --warmup 3 ) > $o/bench_n8.log 2>&1; grep "^{" $o/bench_n8.log | cut -c1-600
