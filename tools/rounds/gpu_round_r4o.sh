#!/bin/bash
# GPU round r4o: CTA-pair GEMMs (cta_group::2) at 32768 tokens -- never measured at this size
for o in gemm_pair=0 gemm_pair=1 gemm_pair=2 gemm_pair=0; do
  echo "== $o"; timeout 120 python tools/encoder_bench.py --shape=128,256 --prec=bf16 --opt=$o 2>&1 | tail -1
done
