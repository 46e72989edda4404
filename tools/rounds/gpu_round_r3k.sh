#!/bin/bash
# GPU round r3k: ncu --set full of the 1 x 1k launch (ot_fused_v7_kernel<768, true>), source page
set -x
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ot_fused_v7 -s 6 -c 1 -o gpurun_out/r3k_lat python tools/quick_bench.py 1000 > gpurun_out/r3k_log.txt 2>&1
ncu -i gpurun_out/r3k_lat.ncu-rep --page raw --csv > gpurun_out/r3k_raw.csv 2>/dev/null
ncu -i gpurun_out/r3k_lat.ncu-rep --page source --csv > gpurun_out/r3k_src.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/r3k_raw.csv
python tools/ncu_src_summary.py gpurun_out/r3k_src.csv ot_fused 30
