#!/bin/bash
# single-sweep kNN (DGR_KNN_SWEEPS=1): bit-identity tests against the fp32 kernel, pipeline tests, bench A/B
set -x
mkdir -p gpurun_out
export DGR_KNN_SWEEPS=1
(timeout 300 python -m pytest tests/test_gpu_knn_registration.py tests/test_gpu_zzz_fullsize.py tests/test_gpu_executor.py -x -q -k "knn or pair_register or batch" 2>&1 | tail -5) > gpurun_out/u_pytest.log 2>&1
(timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/u_bench_1gpu_single_sweep.json) 2> gpurun_out/u_bench_1gpu.err
tail -n 4 gpurun_out/u_pytest.log gpurun_out/u_bench_1gpu.err
