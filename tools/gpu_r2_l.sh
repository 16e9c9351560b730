#!/bin/bash
# round-2 evidence run: the driver's own sequence (suite with -x, smoke, bench) + launch list + ncu --set full + memcheck
set -x
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15) > gpurun_out/l_pytest_all.log 2>&1
(timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5) > gpurun_out/l_smoke.log 2>&1
(timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/l_bench_1gpu.json) 2> gpurun_out/l_bench_1gpu.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/l_launches_native.csv python tools/profile_pair.py > gpurun_out/l_ncu_launches.log 2>&1
(timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_coords.py tests/test_gpu_executor.py -x -q -k "coords or coarse or kernel_map or strided or hash or unique or quantize or conv1_from" 2>&1 | grep -v "Host Frame" | tail -25) > gpurun_out/l_memcheck.log 2>&1
bash tools/gpu_r2_g.sh
cp gpurun_out/g_ncu_pair_raw.csv gpurun_out/l_ncu_pair_raw.csv; cp gpurun_out/g_ncu_misc_raw.csv gpurun_out/l_ncu_misc_raw.csv
ls -la gpurun_out | tail -12; du -sh gpurun_out
