#!/bin/bash
# GPU round P: v7 with parked waits -- parity, micro-bench, judged bench, ubench
set -x
mkdir -p gpurun_out
tools/ubench2 > gpurun_out/p_ubench2.txt 2>&1; cat gpurun_out/p_ubench2.txt
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/p_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/p_pytest.txt
tail -4 gpurun_out/p_pytest.txt | cut -c1-300
timeout 300 python tools/quick_bench.py 1000 64000 256000 512000 1024000 > gpurun_out/p_quick.txt 2>&1
cat gpurun_out/p_quick.txt
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/p_bench.txt 2>&1
tail -1 gpurun_out/p_bench.txt | cut -c1-1800
