#!/bin/bash
set -x
mkdir -p gpurun_out
(timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/m_bench_2gpu.json) 2> gpurun_out/m_bench_2gpu.err
(timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --pairs 32 --warmup 3 > gpurun_out/m_bench_2gpu_strong32.json) 2> gpurun_out/m_bench_2gpu_strong32.err
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/m_bench_2gpu_ref.json) 2> gpurun_out/m_bench_2gpu_ref.err &
REFPID=$!
sleep 20; kill $REFPID 2>/dev/null    # only checks that ranks != 0 exit at once and rank 0 starts; the full-size CPU pair is not waited for
tail -3 gpurun_out/*.err
