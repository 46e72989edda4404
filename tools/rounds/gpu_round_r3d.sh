#!/bin/bash
# GPU round r3d: all-pairs otAspire with a quarter of the exponentials on the FMA/ALU pipes (ex2_poly2) vs all on MUFU
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r3d_ab.txt
timeout 300 python -m pytest tests/test_parity_ot_gpu.py tests/test_kernels_misc_gpu.py -m gpu -q --timeout 120 -k "allpairs or rank_corpus" 2>&1 | tail -2
for v in "" nopoly; do
  if [ -n "$v" ]; then export ASPIRE_B200_LIB=/root/repo/experiments/lib/libaspire_b200_$v.so; fi
  echo "== variant ${v:-intree (poly 1/4)}" >> gpurun_out/r3d_ab.txt
  timeout 300 python tools/side_bench.py otallpairs 2>&1 | head -2 >> gpurun_out/r3d_ab.txt
done
cat gpurun_out/r3d_ab.txt
