#!/bin/bash
set -x
mkdir -p gpurun_out
(timeout 300 python tools/kmap_determinism.py 3 2>&1 | tail -6) > gpurun_out/d_determinism.log 2>&1
(timeout 600 compute-sanitizer --tool memcheck python tools/kmap_determinism.py 3 2>&1 | grep -v "Host Frame" | head -60) > gpurun_out/d_determinism_memcheck.log 2>&1
(timeout 600 python -m pytest tests/test_gpu_executor.py tests/test_gpu_training.py tests/test_gpu_reference_on_shim.py -q -s 2>&1 | tail -60) > gpurun_out/d_pytest_new.log 2>&1
(DGR_BENCH_INFLIGHT=1 timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/d_bench_if1.json) 2> gpurun_out/d_bench_if1.err
(timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/d_bench_if2.json) 2> gpurun_out/d_bench_if2.err
(DGR_TC_F16=0 timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/d_bench_if2_tf32.json) 2> gpurun_out/d_bench_if2_tf32.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/d_launches_native.csv python tools/profile_pair.py > gpurun_out/d_ncu_launches.log 2>&1
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60) > gpurun_out/d_pytest_all.log 2>&1
ls -la gpurun_out; du -sh gpurun_out
