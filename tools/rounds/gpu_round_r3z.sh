#!/bin/bash
# GPU round r3z: pipelined attention after the key-count prefetch: parity + alternating A/B (3 repetitions)
set -x
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_encoder_gpu.py -x -q -m gpu -k "tcgen05" 2>&1 | tail -4
for rep in 1 2 3; do
  for m in 3 5; do
    echo "== attn_tc=$m"
    timeout 120 python tools/encoder_bench.py --shape=128,256 --prec=bf16 --opt=attn_tc=$m 2>&1 | tail -1
  done
done
for m in 3 5; do echo "== attn_tc=$m, small batches"; timeout 120 python tools/encoder_bench.py --shape=8,256 --shape=32,256 --shape=32,128 --shape=64,200 --prec=bf16 --opt=attn_tc=$m 2>&1 | tail -4; done
