#!/bin/bash
# GPU round r2s: var-len kernel ring depth A/B; ncu --set full of the var-len kernel at the bench size (B = 100k), OT and tsAspire modes
set -x
mkdir -p gpurun_out
for v in "" ring36 ring20; do
  if [ -n "$v" ]; then export ASPIRE_B200_LIB=/root/repo/experiments/lib/libaspire_b200_$v.so; fi
  echo "== varlen variant ${v:-intree (ring 28)}" >> gpurun_out/r2s_ab.txt
  timeout 300 python tools/side_bench.py varlen 2>&1 | grep -E "eps=0.1|tsAspire" >> gpurun_out/r2s_ab.txt
done
unset ASPIRE_B200_LIB
cat gpurun_out/r2s_ab.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ot_varlen_kernel -s 4 -c 1 -o gpurun_out/r2s_varlen_ot python tools/side_bench.py varlen > gpurun_out/r2s_ncu_log.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ot_varlen_kernel -s 13 -c 1 -o gpurun_out/r2s_varlen_ts python tools/side_bench.py varlen >> gpurun_out/r2s_ncu_log.txt 2>&1
for m in ot ts; do
ncu -i gpurun_out/r2s_varlen_$m.ncu-rep --page raw --csv > gpurun_out/r2s_raw_$m.csv 2>/dev/null
ncu -i gpurun_out/r2s_varlen_$m.ncu-rep --page source --csv > gpurun_out/r2s_src_$m.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/r2s_raw_$m.csv | head -12
done
