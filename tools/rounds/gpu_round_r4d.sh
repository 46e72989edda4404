#!/bin/bash
# GPU round r4d (2 GPUs): judged bench at N=2, both arms, as the driver launches them; configs[3] shape shortened through rank_corpus_ot
set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r4d_bench_ref_2gpu.txt 2>&1; tail -1 gpurun_out/r4d_bench_ref_2gpu.txt | cut -c1-400
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r4d_bench_2gpu.txt 2>&1; tail -1 gpurun_out/r4d_bench_2gpu.txt | cut -c1-1800
