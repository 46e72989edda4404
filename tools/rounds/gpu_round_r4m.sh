#!/bin/bash
# GPU round r4m: ncu --set full (source page) of the attn-out GEMM (K = 768, residual + LayerNorm-on-read epilogue) at 32768 tokens, final build
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_persistent -s 41 -c 1 -o gpurun_out/r4m_attnout python tools/encoder_bench.py --shape=128,256 --prec=bf16 > gpurun_out/r4m_log.txt 2>&1
ncu -i gpurun_out/r4m_attnout.ncu-rep --page raw --csv > gpurun_out/r4m_raw.csv 2>/dev/null
ncu -i gpurun_out/r4m_attnout.ncu-rep --page source --csv > gpurun_out/r4m_src.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/r4m_raw.csv | head -32
python tools/ncu_src_summary.py gpurun_out/r4m_src.csv gemm 24 | head -70
