#!/bin/bash
# GPU round r3i (2 GPUs): judged bench at N=2 with the final bench.py, reference arm under torchrun
set -x
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r3i_bench_2gpu.txt 2>&1; tail -1 gpurun_out/r3i_bench_2gpu.txt | cut -c1-900
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r3i_bench_ref_2gpu.txt 2>&1; tail -1 gpurun_out/r3i_bench_ref_2gpu.txt | cut -c1-300
timeout 200 python -m pytest tests/test_sharded_gloo.py -q 2>&1 | tail -1
