#!/bin/bash
# GPU round r4h: residual epilogue with hoisted LayerNorm-on-read operands and early residual loads -- parity, A/B, launch list
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_encoder_gpu.py -x -q -m gpu 2>&1 | tail -3
for rep in 1 2; do
  for lib in experiments/lib/libaspire_b200_base.so ""; do
    echo "== lib=${lib:-in-tree (hoisted)}"
    ASPIRE_B200_LIB=$lib timeout 200 python tools/encoder_bench.py --shape=128,256 2>&1 | tail -1
  done
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r4h_enc_launches.csv python tools/encoder_bench.py --shape=128,256 --prec=bf16 > gpurun_out/r4h_enc_log.txt 2>&1
grep "gemm_tn_persistent_kernel<192, 2, 1>" gpurun_out/r4h_enc_launches.csv | tail -4 | awk -F'","' '{print $NF}'
