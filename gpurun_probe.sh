#!/bin/bash
# ad-hoc GPU probe used during development: 4-GPU weak scaling line
mkdir -p gpurun_out/r01i
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 2 --warmup 3 ) > gpurun_out/r01i/bench_n4.log 2>&1
tail -3 gpurun_out/r01i/bench_n4.log | cut -c1-400
