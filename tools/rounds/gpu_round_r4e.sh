#!/bin/bash
# GPU round r4e: GEMM bf16 epilogue that does not compute the lo halves when there is no lo output -- A/B
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_encoder_gpu.py -x -q -m gpu 2>&1 | tail -3
for rep in 1 2 3; do
  for lib in experiments/lib/libaspire_b200_base.so ""; do
    echo "== lib=${lib:-in-tree (hi only)}"
    ASPIRE_B200_LIB=$lib timeout 200 python tools/encoder_bench.py --shape=128,256 --prec=bf16 2>&1 | tail -1
  done
done
