#!/bin/bash
# GPU round O: v7 fused kernel (8 Gram + 4 Sinkhorn warps) -- parity, micro-bench, ncu capture
set -x
mkdir -p gpuruo_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 > gpuruo_out/o_pytest.txt 2>&1; echo "pytest exit $?" >> gpuruo_out/o_pytest.txt
tail -5 gpuruo_out/o_pytest.txt | cut -c1-300
timeout 300 python tools/quick_bench.py 1000 64000 256000 1024000 > gpuruo_out/o_quick.txt 2>&1
cat gpuruo_out/o_quick.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ot_fused -s 4 -c 1 -o gpuruo_out/o_fused python tools/quick_bench.py 256000 > gpuruo_out/o_ncu.log 2>&1
tail -3 gpuruo_out/o_ncu.log
