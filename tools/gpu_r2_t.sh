#!/bin/bash
# after reverting the four-chain fill: executor + golden tests, the driver's 1-GPU bench line
set -x
mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_gpu_executor.py tests/test_gpu_zzzz_golden_fullsize.py -x -q 2>&1 | tail -4) > gpurun_out/t_pytest.log 2>&1
(timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/t_bench_1gpu.json) 2> gpurun_out/t_bench_1gpu.err
tail -n 3 gpurun_out/t_pytest.log gpurun_out/t_bench_1gpu.err
