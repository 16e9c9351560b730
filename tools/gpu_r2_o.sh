#!/bin/bash
set -x
mkdir -p gpurun_out
(DGR_BENCH_INFLIGHT=1 timeout 400 python bench.py --steps 12 --warmup 3 > gpurun_out/o_bench_nc.json) 2> gpurun_out/o_bench_nc.err
DGR_EXTRA_NVCC_FLAGS=-DDGR_GATHER_CG python -m deepglobalregistration_b200.build --force > gpurun_out/o_build.log 2>&1
(DGR_BENCH_INFLIGHT=1 timeout 400 python bench.py --steps 12 --warmup 3 > gpurun_out/o_bench_cg.json) 2> gpurun_out/o_bench_cg.err
(timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/o_bench_cg_if4.json) 2> gpurun_out/o_bench_cg_if4.err
ls -la gpurun_out | tail -4
