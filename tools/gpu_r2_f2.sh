#!/bin/bash
set -x
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_executor.py -q 2>&1 | tail -30) > gpurun_out/f2_pytest_exec.log 2>&1
(timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/f2_bench_if2.json) 2> gpurun_out/f2_bench_if2.err
(DGR_BENCH_INFLIGHT=3 timeout 400 python bench.py --steps 21 --warmup 5 > gpurun_out/f2_bench_if3.json) 2> gpurun_out/f2_bench_if3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/f2_launches_native.csv python tools/profile_pair.py > gpurun_out/f2_ncu_launches.log 2>&1
(timeout 900 python -m pytest tests/test_gpu_zzzz_golden_fullsize.py tests/test_gpu_pipeline.py tests/test_gpu_zzz_fullsize.py -q 2>&1 | tail -30) > gpurun_out/f2_pytest_big.log 2>&1
bash tools/gpu_r2_g.sh
