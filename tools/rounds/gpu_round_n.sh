#!/bin/bash
# GPU round N: v7 fused kernel (8 Gram + 4 Sinkhorn warps) -- parity, micro-bench, ncu capture
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/n_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/n_pytest.txt
tail -5 gpurun_out/n_pytest.txt | cut -c1-300
timeout 300 python tools/quick_bench.py 1000 64000 256000 1024000 > gpurun_out/n_quick.txt 2>&1
cat gpurun_out/n_quick.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ot_fused -s 4 -c 1 -o gpurun_out/n_fused python tools/quick_bench.py 256000 > gpurun_out/n_ncu.log 2>&1
tail -3 gpurun_out/n_ncu.log
