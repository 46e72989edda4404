#!/bin/bash
# GPU round r3o: ncu --set full (source page) of the four GEMMs of one encoder layer at B=128 L=256 -- what the epilogue warps wait on
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_persistent -s 40 -c 4 -o gpurun_out/r3o_gemm python tools/encoder_bench.py --shape=128,256 --prec=bf16 > gpurun_out/r3o_log.txt 2>&1
ncu -i gpurun_out/r3o_gemm.ncu-rep --page raw --csv > gpurun_out/r3o_raw.csv 2>/dev/null
ncu -i gpurun_out/r3o_gemm.ncu-rep --page source --csv > gpurun_out/r3o_src.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/r3o_raw.csv | grep -v "^  l1tex\|lts__" 
python tools/ncu_src_summary.py gpurun_out/r3o_src.csv gemm 22
