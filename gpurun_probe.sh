#!/bin/bash
mkdir -p gpurun_out/r01d
nvidia-smi -L
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 ) > gpurun_out/r01d/bench_n2.log 2>&1
tail -3 gpurun_out/r01d/bench_n2.log | cut -c1-900
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --ref-cells 20000 ) > gpurun_out/r01d/bench_ref_n2.log 2>&1
tail -2 gpurun_out/r01d/bench_ref_n2.log | cut -c1-300
