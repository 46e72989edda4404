#!/bin/bash
# GPU round r3f: span pooling staged by cp.async.bulk vs streaming loads
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r3f_ab.txt
timeout 300 python -m pytest tests/test_kernels_misc_gpu.py tests/test_encoder_gpu.py -m gpu -q --timeout 120 -k "span_pool or readme or entity" 2>&1 | tail -2
for m in 1 0; do
  echo "== span_tma=$m" >> gpurun_out/r3f_ab.txt
  ASP_OPTIONS=span_tma=$m timeout 200 python tools/side_bench.py pool >> gpurun_out/r3f_ab.txt 2>&1
done
cat gpurun_out/r3f_ab.txt
ASP_OPTIONS=span_tma=1 timeout 300 ncu --set full --clock-control none -k regex:span_mean_pool_tma -s 2 -c 1 -o gpurun_out/r3f_span python tools/side_bench.py pool > gpurun_out/r3f_ncu_log.txt 2>&1
ncu -i gpurun_out/r3f_span.ncu-rep --page raw --csv > gpurun_out/r3f_raw.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/r3f_raw.csv | head -12
