#!/bin/bash
# GPU round r3v: attention with P kept in tensor memory (attn_tc=4): parity + encoder A/B
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_encoder_gpu.py -x -q -m gpu -k tcgen05 2>&1 | tail -8
for rep in 1 2; do
  for m in 3 4; do
    echo "== attn_tc=$m"
    timeout 200 python tools/encoder_bench.py --shape=128,256 --prec=bf16 --opt=attn_tc=$m 2>&1 | tail -1
  done
done
