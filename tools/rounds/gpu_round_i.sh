#!/bin/bash
# GPU round I: judged bench (N=1), reference arm, ncu launch list of the bench command
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/i_bench.txt 2>&1
tail -1 gpurun_out/i_bench.txt | cut -c1-3000
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/i_bench_ref.txt 2>&1
tail -1 gpurun_out/i_bench_ref.txt | cut -c1-1200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/i_launches.csv python bench.py --steps 5 --warmup 3 --cpu-seconds 1 > gpurun_out/i_ncu_bench.log 2>&1
grep -c ot_fused gpurun_out/i_launches.csv
