#!/bin/bash
# GPU round r3l: per-step log2e/eps table in the fused solvers -- parity + A/B (1 x 1k latency, 256k burst, sustained)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_ot_gpu.py tests/test_parity_ot_spec_gpu.py -x -q -m gpu 2>&1 | tail -4
for rep in 1 2; do
  for lib in "" experiments/lib/libaspire_b200_nott.so; do
    echo "== lib=${lib:-in-tree}"
    ASPIRE_B200_LIB=$lib timeout 200 python tools/quick_bench.py 1000 256000 2>&1 | tail -4
  done
done
for lib in "" experiments/lib/libaspire_b200_nott.so; do
  ASPIRE_B200_LIB=$lib ASP_STEPS=500 timeout 200 python tools/sustained_ab.py 2>&1 | tail -2
done
