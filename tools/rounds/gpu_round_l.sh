#!/bin/bash
# GPU round L: side kernels (all-pairs tsAspire on tcgen05, var-length OT, span pool, top-k, encoder) + ncu captures
set -x
mkdir -p gpurun_out
timeout 600 python tools/side_bench.py > gpurun_out/l_side.txt 2>&1; cat gpurun_out/l_side.txt
timeout 600 python tools/encoder_bench.py > gpurun_out/l_encoder.txt 2>&1; cat gpurun_out/l_encoder.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:l2max_allpairs -s 2 -c 1 -o gpurun_out/l_allpairs python tools/side_bench.py allpairs > gpurun_out/l_ncu1.log 2>&1; tail -2 gpurun_out/l_ncu1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"sinkhorn_warp|pair_cost" -s 2 -c 2 -o gpurun_out/l_varlen python tools/side_bench.py varlen > gpurun_out/l_ncu2.log 2>&1; tail -2 gpurun_out/l_ncu2.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"span_mean_pool|topk_kernel" -s 2 -c 2 -o gpurun_out/l_pool_topk python tools/side_bench.py pool topk > gpurun_out/l_ncu3.log 2>&1; tail -2 gpurun_out/l_ncu3.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_tn|attention_kernel" -s 40 -c 4 -o gpurun_out/l_encoder python tools/encoder_bench.py --quick > gpurun_out/l_ncu4.log 2>&1; tail -2 gpurun_out/l_ncu4.log
