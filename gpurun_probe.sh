#!/bin/bash
# ad-hoc GPU probe used during development: 2-GPU weak scaling line
mkdir -p gpurun_out/r01i
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 2 --warmup 3 ) > gpurun_out/r01i/bench_n2.log 2>&1
tail -1 gpurun_out/r01i/bench_n2.log | cut -c1-200
