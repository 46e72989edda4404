#!/bin/bash
# GPU round r2w: headline FFMA2 kernel, bytes-in-flight A/B (cp.async ring depth vs cost-tile slots), sustained
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2w_ab.txt
for v in "" r4s5 r4s4; do
  if [ -n "$v" ]; then export ASPIRE_B200_LIB=/root/repo/experiments/lib/libaspire_b200_$v.so; fi
  echo "== variant ${v:-intree (ring 5, slots 4)}" >> gpurun_out/r2w_ab.txt
  ASP_STEPS=500 timeout 120 python tools/sustained_ab.py >> gpurun_out/r2w_ab.txt 2>&1
done
unset ASPIRE_B200_LIB
cat gpurun_out/r2w_ab.txt
