#!/bin/bash
# GPU round r3s: ncu --set full (source page) of attention_tc_kernel at B=128 L=256
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 11 -c 1 -o gpurun_out/r3s_attn python tools/encoder_bench.py --shape=128,256 --prec=bf16 > gpurun_out/r3s_log.txt 2>&1
ncu -i gpurun_out/r3s_attn.ncu-rep --page raw --csv > gpurun_out/r3s_raw.csv 2>/dev/null
ncu -i gpurun_out/r3s_attn.ncu-rep --page source --csv > gpurun_out/r3s_src.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/r3s_raw.csv
python tools/ncu_src_summary.py gpurun_out/r3s_src.csv attention 40
