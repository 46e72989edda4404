#!/bin/bash
# GPU round r3p: 12/16 epilogue warps in the persistent GEMM (64 columns per warp) vs 8; swizzled 64-byte staging
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder_gpu.py -x -q -m gpu 2>&1 | tail -3
for rep in 1 2; do
  for lib in "" experiments/lib/libaspire_b200_epi8.so experiments/lib/libaspire_b200_episingle.so; do
    echo "== lib=${lib:-in-tree}"
    ASPIRE_B200_LIB=$lib timeout 200 python tools/encoder_bench.py --shape=128,256 --shape=32,256 2>&1 | tail -2
  done
done
ASPIRE_B200_LIB=experiments/lib/libaspire_b200_episingle.so timeout 600 python -m pytest tests/test_encoder_gpu.py -x -q -m gpu 2>&1 | tail -3
