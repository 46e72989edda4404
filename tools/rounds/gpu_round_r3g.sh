#!/bin/bash
# GPU round r3g: ncu --set full of topk_stream_kernel (1k x 125k)
set -x
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:topk_stream -s 2 -c 1 -o gpurun_out/r3g_topk python tools/side_bench.py topk > gpurun_out/r3g_log.txt 2>&1
ncu -i gpurun_out/r3g_topk.ncu-rep --page raw --csv > gpurun_out/r3g_raw.csv 2>/dev/null
ncu -i gpurun_out/r3g_topk.ncu-rep --page source --csv > gpurun_out/r3g_src.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/r3g_raw.csv
python tools/ncu_src_summary.py gpurun_out/r3g_src.csv topk_stream 22
