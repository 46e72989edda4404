#!/bin/bash
# GPU round r3b: ncu --set full of the tcgen05 attention kernel (B=32, L=256)
set -x
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 30 -c 1 -o gpurun_out/r3b_attn python tools/encoder_bench.py --quick > gpurun_out/r3b_log.txt 2>&1
ncu -i gpurun_out/r3b_attn.ncu-rep --page raw --csv > gpurun_out/r3b_raw.csv 2>/dev/null
ncu -i gpurun_out/r3b_attn.ncu-rep --page source --csv > gpurun_out/r3b_src.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/r3b_raw.csv
python tools/ncu_src_summary.py gpurun_out/r3b_src.csv attention_tc 25
