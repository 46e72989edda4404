#!/bin/bash
# GPU round r3y: ncu --set full (source page) of attention_tc_pipe_kernel at B=128 L=256
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc_pipe -s 11 -c 1 -o gpurun_out/r3y_attn python tools/encoder_bench.py --shape=128,256 --prec=bf16 --opt=attn_tc=5 > gpurun_out/r3y_log.txt 2>&1
ncu -i gpurun_out/r3y_attn.ncu-rep --page raw --csv > gpurun_out/r3y_raw.csv 2>/dev/null
ncu -i gpurun_out/r3y_attn.ncu-rep --page source --csv > gpurun_out/r3y_src.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/r3y_raw.csv
python tools/ncu_src_summary.py gpurun_out/r3y_src.csv attention 45
