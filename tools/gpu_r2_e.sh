#!/bin/bash
set -x
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_executor.py tests/test_gpu_reference_on_shim.py -q -x 2>&1 | tail -30) > gpurun_out/e_pytest_new.log 2>&1
(timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_executor.py -x -q -k "coarse or kernel_map or strided" 2>&1 | grep -v "Host Frame" | tail -25) > gpurun_out/e_memcheck.log 2>&1
(DGR_BENCH_INFLIGHT=1 timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/e_bench_if1.json) 2> gpurun_out/e_bench_if1.err
(timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/e_bench_if2.json) 2> gpurun_out/e_bench_if2.err
(DGR_BENCH_INFLIGHT=3 timeout 400 python bench.py --steps 21 --warmup 5 > gpurun_out/e_bench_if3.json) 2> gpurun_out/e_bench_if3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/e_launches_native.csv python tools/profile_pair.py > gpurun_out/e_ncu_launches.log 2>&1
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40) > gpurun_out/e_pytest_all.log 2>&1
ls -la gpurun_out | tail -12; du -sh gpurun_out
