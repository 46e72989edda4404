#!/bin/bash
# GPU round r2q (8 GPUs): BASELINE configs[3] at full size through the all-pairs kernel; judged bench at N=8 (packed gather on a side stream)
set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/config4.py > gpurun_out/r2q_config4_8gpu.txt 2>&1; tail -1 gpurun_out/r2q_config4_8gpu.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2q_bench_8gpu.txt 2>&1; tail -1 gpurun_out/r2q_bench_8gpu.txt | cut -c1-1500
