#!/bin/bash
set -x
mkdir -p gpurun_out
(timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/p_bench_8gpu.json) 2> gpurun_out/p_bench_8gpu.err
tail -n 4 gpurun_out/p_bench_8gpu.err
