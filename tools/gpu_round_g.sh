#!/bin/bash
# GPU round G: staggered 8-warp fused kernel -- parity, micro-bench, ncu capture
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/g_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/g_pytest.txt
tail -5 gpurun_out/g_pytest.txt | cut -c1-300
timeout 300 python tools/quick_bench.py 1000 8000 64000 256000 1024000 > gpurun_out/g_quick.txt 2>&1
cat gpurun_out/g_quick.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ot_fused -s 4 -c 1 -o gpurun_out/g_fused python tools/quick_bench.py 64000 > gpurun_out/g_ncu.log 2>&1
tail -3 gpurun_out/g_ncu.log
