#!/bin/bash
# GPU round r2z: streaming top-k, chunks per CTA A/B (1 / 4 / 16) with per-kernel durations
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2z_ab.txt
timeout 300 python -m pytest tests/test_kernels_misc_gpu.py -m gpu -q --timeout 300 -k "topk or shard" 2>&1 | tail -1
for v in "" span1 span16; do
  if [ -n "$v" ]; then export ASPIRE_B200_LIB=/root/repo/experiments/lib/libaspire_b200_$v.so; fi
  echo "== variant ${v:-intree (span 4)}" >> gpurun_out/r2z_ab.txt
  timeout 200 python tools/side_bench.py topk >> gpurun_out/r2z_ab.txt 2>&1
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:topk -c 3 --csv --log-file gpurun_out/r2z_$v.csv python tools/side_bench.py topk > /dev/null 2>&1
  grep -E "topk" gpurun_out/r2z_$v.csv | awk -F'","' '{print substr($5,1,24), $NF}' >> gpurun_out/r2z_ab.txt
done
cat gpurun_out/r2z_ab.txt
