#!/bin/bash
# round-2 GPU call B: the native executor - unit tests first (memcheck on the small ones), then the suite, then bench
set -x
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_executor.py -x -q -s 2>&1 | tail -40) > gpurun_out/b_pytest_exec.log 2>&1
(timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_executor.py -x -q -k "coarse or kernel_map or strided" 2>&1 | tail -30) > gpurun_out/b_memcheck.log 2>&1
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/b_pytest_all.log 2>&1
(timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/b_bench_if2.json) 2> gpurun_out/b_bench_if2.err
(DGR_BENCH_INFLIGHT=1 timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/b_bench_if1.json) 2> gpurun_out/b_bench_if1.err
(DGR_BENCH_INFLIGHT=3 timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/b_bench_if3.json) 2> gpurun_out/b_bench_if3.err
ls -la gpurun_out
