#!/bin/bash
# GPU round r2r: sleep-polling A/B -- var-len kernel (idle Sinkhorn / producer waits) and all-pairs OT kernel (front warps)
set -x
mkdir -p gpurun_out
for v in "" s200 s200p100 s500p200 s100p50; do
  if [ -n "$v" ]; then export ASPIRE_B200_LIB=/root/repo/experiments/lib/libaspire_b200_$v.so; fi
  echo "== varlen variant ${v:-intree}" >> gpurun_out/r2r_ab.txt
  timeout 300 python tools/side_bench.py varlen >> gpurun_out/r2r_ab.txt 2>&1
done
unset ASPIRE_B200_LIB
for v in "" oa0; do
  if [ -n "$v" ]; then export ASPIRE_B200_LIB=/root/repo/experiments/lib/libaspire_b200_$v.so; fi
  echo "== otallpairs variant ${v:-intree (front 400 ns, drain 200 ns)}" >> gpurun_out/r2r_ab.txt
  timeout 300 python tools/side_bench.py otallpairs 2>&1 | head -1 >> gpurun_out/r2r_ab.txt
done
unset ASPIRE_B200_LIB
cat gpurun_out/r2r_ab.txt
timeout 300 python -m pytest tests/test_parity_ot_gpu.py tests/test_varlen_gpu.py -m gpu -q --timeout 120 2>&1 | tail -2
