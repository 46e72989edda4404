#!/bin/bash
# GPU round AA: final ncu captures of the side kernels after their round-1 changes
set -x
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none -k regex:l2max_allpairs -s 2 -c 1 -o gpurun_out/aa_allpairs python tools/side_bench.py allpairs > gpurun_out/aa_ncu1.log 2>&1; tail -1 gpurun_out/aa_ncu1.log
timeout 300 ncu --set full --clock-control none -k regex:"sinkhorn_warp|pair_cost_cta" -s 2 -c 2 -o gpurun_out/aa_varlen python tools/side_bench.py varlen > gpurun_out/aa_ncu2.log 2>&1; tail -1 gpurun_out/aa_ncu2.log
timeout 300 python tools/side_bench.py > gpurun_out/aa_side.txt 2>&1; cat gpurun_out/aa_side.txt
