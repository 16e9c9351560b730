#!/bin/bash
# final confirmation of the round-2 tree: whole GPU suite, smoke(), the driver's 1-GPU bench line
set -x
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6) > gpurun_out/s_pytest_all.log 2>&1
(timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3) > gpurun_out/s_smoke.log 2>&1
(timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/s_bench_1gpu.json) 2> gpurun_out/s_bench_1gpu.err
tail -n 3 gpurun_out/s_pytest_all.log gpurun_out/s_smoke.log gpurun_out/s_bench_1gpu.err
