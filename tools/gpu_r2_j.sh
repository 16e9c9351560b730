#!/bin/bash
set -x
mkdir -p gpurun_out
(timeout 400 python bench.py --steps 21 --warmup 5 > gpurun_out/j_bench_if3.json) 2> gpurun_out/j_bench_if3.err
(DGR_TC_PREFETCH=2 timeout 400 python bench.py --steps 21 --warmup 5 > gpurun_out/j_bench_if3_pd2.json) 2> gpurun_out/j_bench_if3_pd2.err
(DGR_BENCH_INFLIGHT=4 timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/j_bench_if4.json) 2> gpurun_out/j_bench_if4.err
(timeout 300 python -m pytest tests/test_gpu_executor.py -q -k "output_stationary or pair_register or net_forward" 2>&1 | tail -8) > gpurun_out/j_pytest.log 2>&1
ls -la gpurun_out | tail -5
