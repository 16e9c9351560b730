#!/bin/bash
set -x
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_executor.py -q -x 2>&1 | tail -30) > gpurun_out/f_pytest_exec.log 2>&1
(timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/f_bench_if2.json) 2> gpurun_out/f_bench_if2.err
(DGR_TC_OS=0 timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/f_bench_if2_noos.json) 2> gpurun_out/f_bench_if2_noos.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/f_launches_native.csv python tools/profile_pair.py > gpurun_out/f_ncu_launches.log 2>&1
(timeout 900 python -m pytest tests/test_gpu_zzzz_golden_fullsize.py tests/test_gpu_pipeline.py tests/test_gpu_zzz_fullsize.py -q 2>&1 | tail -30) > gpurun_out/f_pytest_big.log 2>&1
ls -la gpurun_out | tail -8
