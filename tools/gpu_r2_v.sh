#!/bin/bash
# default (two-sweep) kNN path after the single-sweep variant was added next to it
set -x
mkdir -p gpurun_out
(timeout 120 python -m pytest tests/test_gpu_knn_registration.py tests/test_gpu_executor.py -x -q -k "knn or pair_register" 2>&1 | tail -3) > gpurun_out/v_pytest.log 2>&1
(timeout 100 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/v_bench_1gpu.json) 2> gpurun_out/v_bench_1gpu.err
tail -n 3 gpurun_out/v_pytest.log gpurun_out/v_bench_1gpu.err
