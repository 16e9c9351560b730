#!/bin/bash
set -x
mkdir -p gpurun_out
(timeout 600 python tools/run_configs.py > gpurun_out/n_configs.log) 2> gpurun_out/n_configs.err
(timeout 400 python bench.py --pairs 12 --warmup 3 > gpurun_out/n_bench_strong12.json) 2> gpurun_out/n_bench_strong12.err
tail -3 gpurun_out/n_*.err
