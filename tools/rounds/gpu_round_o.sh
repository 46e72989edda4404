#!/bin/bash
# GPU round O2: v7 fused kernel -- ncu capture at 256k pairs
set -x
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ot_fused_v7 -s 4 -c 1 -o gpurun_out/o_fused python tools/quick_bench.py 256000 > gpurun_out/o_ncu.log 2>&1
tail -3 gpurun_out/o_ncu.log
