#!/bin/bash
# GPU round r4c: L2 prefetch of the next tile's A row block in the persistent GEMM: off / all CTAs / column-tile-0 CTAs only
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_encoder_gpu.py -x -q -m gpu 2>&1 | tail -3
for rep in 1 2; do
  for lib in experiments/lib/libaspire_b200_apf0.so "" experiments/lib/libaspire_b200_apf1.so; do
    echo "== lib=${lib:-in-tree (mode 2)}"
    ASPIRE_B200_LIB=$lib timeout 200 python tools/encoder_bench.py --shape=128,256 2>&1 | tail -1
  done
done
