#!/bin/bash
# GPU round r4q: FFN2 (residual epilogue, K = 3072) on CTA pairs too: parity + A/B against the previous build
timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_encoder_gpu.py -x -q -m gpu 2>&1 | grep -E "passed|failed|Error" | head -5
for lib in experiments/lib/libaspire_b200_base.so "" experiments/lib/libaspire_b200_base.so ""; do
  echo "== lib=${lib:-in-tree (FFN2 on pairs)}"; ASPIRE_B200_LIB=$lib timeout 120 python tools/encoder_bench.py --shape=128,256 --prec=bf16 2>&1 | tail -1
done
