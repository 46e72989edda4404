#!/bin/bash
# GPU round r3r: quad-friendly lane map in the var-len Gram -- parity + A/B at B=100k
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_varlen_gpu.py tests/test_parity_l2max_gpu.py -x -q -m gpu 2>&1 | tail -3
for rep in 1 2; do
  for lib in "" experiments/lib/libaspire_b200_oldmap.so; do
    echo "== lib=${lib:-in-tree}"
    ASPIRE_B200_LIB=$lib timeout 300 python tools/side_bench.py varlen 2>&1 | tail -4
  done
done
