#!/bin/bash
set -x
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_executor.py tests/test_gpu_knn_registration.py -q 2>&1 | tail -30) > gpurun_out/h_pytest_exec.log 2>&1
(timeout 400 python bench.py --steps 21 --warmup 5 > gpurun_out/h_bench_if3.json) 2> gpurun_out/h_bench_if3.err
(DGR_KNN_COARSE=1 timeout 400 python bench.py --steps 21 --warmup 5 > gpurun_out/h_bench_if3_knncoarse.json) 2> gpurun_out/h_bench_if3_knncoarse.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/h_launches_native.csv python tools/profile_pair.py > gpurun_out/h_ncu_launches.log 2>&1
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/h_pytest_all.log 2>&1
ls -la gpurun_out | tail -6
