#!/bin/bash
# GPU round r2p: rank_corpus_ot test, config 4 shard (1k x 125k) on one GPU, both paths
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_misc_gpu.py tests/test_parity_ot_gpu.py -m gpu -q --timeout 120 -k "rank_corpus or allpairs or topk" 2>&1 | tail -3
timeout 300 python tools/config4.py --candidates 125000 > gpurun_out/r2p_config4_1gpu.txt 2>&1; tail -1 gpurun_out/r2p_config4_1gpu.txt
timeout 300 python tools/config4.py --candidates 125000 --per-query >> gpurun_out/r2p_config4_1gpu.txt 2>&1; tail -1 gpurun_out/r2p_config4_1gpu.txt
