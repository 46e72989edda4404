#!/bin/bash
# GPU round r2o: all-pairs otAspire, 8 vs 12 Sinkhorn warps; ncu --set full of the 8-warp kernel on a 1k x 4k problem
set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_parity_ot_gpu.py -m gpu -q --timeout 120 -k "allpairs" 2>&1 | tail -2
for w in 8 12; do
  ASP_OPTIONS=oa_warps=$w timeout 300 python tools/side_bench.py otallpairs 2>&1 | tail -2 | sed "s/^/oa_warps=$w: /" >> gpurun_out/r2o_side.txt
done
ASP_OPTIONS=oa_warps=12 timeout 200 python -m pytest tests/test_parity_ot_gpu.py -m gpu -q --timeout 120 -k "allpairs" 2>&1 | tail -2
cat gpurun_out/r2o_side.txt
ASP_OTAP_NC=4000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:ot_allpairs_kernel -c 1 -o gpurun_out/r2o_otallpairs python tools/side_bench.py otallpairs > gpurun_out/r2o_ncu_log.txt 2>&1
ncu -i gpurun_out/r2o_otallpairs.ncu-rep --page raw --csv > gpurun_out/r2o_raw.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/r2o_raw.csv | tail -32
