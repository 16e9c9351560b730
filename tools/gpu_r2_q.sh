#!/bin/bash
set -x
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6) > gpurun_out/q_pytest_all.log 2>&1
(timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/q_bench_1gpu.json) 2> gpurun_out/q_bench_1gpu.err
tail -n 3 gpurun_out/q_pytest_all.log gpurun_out/q_bench_1gpu.err
