#!/bin/bash
# GPU round r3x: pipelined attention (attn_tc=5): parity (incl. the exact-softmax path) + encoder A/B
set -x
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_encoder_gpu.py -x -q -m gpu -k "tcgen05" 2>&1 | tail -12
for rep in 1 2; do
  for m in 3 5; do
    echo "== attn_tc=$m"
    timeout 120 python tools/encoder_bench.py --shape=128,256 --prec=bf16 --opt=attn_tc=$m 2>&1 | tail -1
  done
done
