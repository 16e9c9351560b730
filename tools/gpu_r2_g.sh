#!/bin/bash
# ncu --set full evidence for every kernel family of the path (raw CSV pages only: the .ncu-rep files stay on the box)
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --profile-from-start off \
  -k 'regex:spconv_tc_pair|spconv_os|spconv_tc_kernel|spconv_table|kmap_probe|kmap_fill|knn_tc_kernel|se3_register|insert_min|unique_scatter|coarse_' \
  -o /tmp/g_ncu_pair python tools/profile_pair.py > gpurun_out/g_ncu_pair.log 2>&1
ncu -i /tmp/g_ncu_pair.ncu-rep --page raw --csv > gpurun_out/g_ncu_pair_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --profile-from-start off -k 'regex:icp_match|ransac' -c 6 \
  -o /tmp/g_ncu_misc python tools/profile_misc.py > gpurun_out/g_ncu_misc.log 2>&1
ncu -i /tmp/g_ncu_misc.ncu-rep --page raw --csv > gpurun_out/g_ncu_misc_raw.csv 2>/dev/null
ls -la /tmp/*.ncu-rep gpurun_out | tail -12; du -sh gpurun_out
