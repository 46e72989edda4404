#!/bin/bash
# round 2, call K: A/B of the Gram unroll policy (steps per trip for small / large tiles): 8/4 (in-tree), 4/4, 4/2, 8/8
set -x
mkdir -p gpurun_out
for v in "" u44 u42 u88; do
  if [ -n "$v" ]; then export ASPIRE_B200_LIB=$PWD/experiments/lib/libaspire_b200_$v.so; fi
  echo "== variant ${v:-intree}" >> gpurun_out/r2k_ab.txt
  timeout 300 python tools/side_bench.py varlen >> gpurun_out/r2k_ab.txt 2>&1
done
cat gpurun_out/r2k_ab.txt
