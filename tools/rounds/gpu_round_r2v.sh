#!/bin/bash
# GPU round r2v: tensor-core pool kernel, fence placement A/B (0 writer-side, 1 consumer-side + piecewise prefetch, 2 none = upper bound, unsafe)
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2v_ab.txt
for v in "" tcf1 tcf2; do
  if [ -n "$v" ]; then export ASPIRE_B200_LIB=/root/repo/experiments/lib/libaspire_b200_$v.so; fi
  echo "== variant ${v:-intree (fence 0)}" >> gpurun_out/r2v_ab.txt
  timeout 120 python -m pytest tests/test_parity_ot_gpu.py -m gpu -q --timeout 60 -x -k "pool_kernel" 2>&1 | tail -1 >> gpurun_out/r2v_ab.txt
  ASP_TC=1 ASP_STEPS=300 timeout 120 python tools/sustained_ab.py >> gpurun_out/r2v_ab.txt 2>&1
done
cat gpurun_out/r2v_ab.txt
