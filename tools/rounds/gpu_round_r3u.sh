#!/bin/bash
# GPU round r3u: LayerNorm on read -- parity (bit identity, HF tolerance) + encoder A/B, with attn_tc 1 and 3
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder_gpu.py tests/test_gemm_gpu.py -x -q -m gpu 2>&1 | tail -8
for rep in 1 2; do
  for o in "ln_on_read=0" "ln_on_read=1"; do
    echo "== $o"
    timeout 200 python tools/encoder_bench.py --shape=128,256 --opt=$o 2>&1 | tail -1
  done
done
echo "== ln_on_read=1 attn_tc=3"
timeout 200 python tools/encoder_bench.py --shape=128,256 --shape=32,256 --shape=8,256 --opt=attn_tc=3 2>&1 | tail -3
