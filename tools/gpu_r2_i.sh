#!/bin/bash
set -x
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_executor.py tests/test_gpu_knn_registration.py tests/test_gpu_pipeline.py tests/test_gpu_reference_on_shim.py -q 2>&1 | tail -15) > gpurun_out/i_pytest.log 2>&1
(timeout 400 python bench.py --steps 21 --warmup 5 > gpurun_out/i_bench_if3.json) 2> gpurun_out/i_bench_if3.err
(timeout 500 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/i_bench_reference.json) 2> gpurun_out/i_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/i_launches_native.csv python tools/profile_pair.py > gpurun_out/i_ncu_launches.log 2>&1
ls -la gpurun_out | tail -6
