#!/bin/bash
# GPU round r3n: L2 bulk prefetch of the residual rows in the persistent GEMM (A/B on the encoder), encoder tests
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder_gpu.py -x -q -m gpu 2>&1 | tail -3
for rep in 1 2; do
  for lib in "" experiments/lib/libaspire_b200_nopf.so; do
    echo "== lib=${lib:-in-tree}"
    ASPIRE_B200_LIB=$lib timeout 200 python tools/encoder_bench.py --shape=128,256 --shape=32,256 2>&1 | tail -2
  done
done
