#!/bin/bash
# GPU round r4k (8 GPUs): judged bench at N=8 on the final build
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 --no-side > gpurun_out/r4k_bench_8gpu.txt 2>&1; tail -1 gpurun_out/r4k_bench_8gpu.txt | cut -c1-1400
