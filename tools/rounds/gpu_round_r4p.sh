#!/bin/bash
# GPU round r4p: auto rule for CTA-pair GEMMs (bf16-output GEMMs of >= 16384 rows): parity + A/B
timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_encoder_gpu.py -x -q -m gpu 2>&1 | grep -E "passed|failed|Error" | head -5
for o in gemm_pair=0 gemm_pair=-1 gemm_pair=0 gemm_pair=-1; do
  echo "== $o"; timeout 120 python tools/encoder_bench.py --shape=128,256 --opt=$o 2>&1 | tail -1
done
