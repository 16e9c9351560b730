#!/bin/bash
# round-2 GPU call A: whole GPU suite with the cta_group::2 conv (variant 3), bench with both variants,
# ncu --set full of the kernels that had no ncu section in round 1.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt
(DGR_TC_VARIANT=3 timeout 1200 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60) > gpurun_out/a_pytest_v3.log 2>&1
(timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench_v1.json) 2> gpurun_out/a_bench_v1.err
(DGR_TC_VARIANT=3 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench_v3.json) 2> gpurun_out/a_bench_v3.err
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k 'regex:kernel_map_table|insert_min|unique_scatter|kernel_map_fill|knn_tc|se3_register|spconv_table|scan_blocks' \
  -o gpurun_out/a_ncu_misc python tools/profile_pair.py > gpurun_out/a_ncu_misc.log 2>&1
ls -la gpurun_out
