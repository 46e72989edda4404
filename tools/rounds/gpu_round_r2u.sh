#!/bin/bash
# GPU round r2u: ncu --set full of the tensor-core pool kernel at the bench size
set -x
mkdir -p gpurun_out
ASP_TC=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:ot_fused_tc_kernel -s 5 -c 1 -o gpurun_out/r2u_tc python tools/sustained_ab.py > gpurun_out/r2u_ncu_log.txt 2>&1
ncu -i gpurun_out/r2u_tc.ncu-rep --page raw --csv > gpurun_out/r2u_raw.csv 2>/dev/null
ncu -i gpurun_out/r2u_tc.ncu-rep --page source --csv > gpurun_out/r2u_src.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/r2u_raw.csv
