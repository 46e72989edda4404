#!/bin/bash
# GPU round R: all-pairs kernel with single-load K-block stages (64B swizzle), heads kernel -- parity + side bench
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/r_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r_pytest.txt
grep -E "FAIL|passed|failed|exit|Error" gpurun_out/r_pytest.txt | cut -c1-250 | tail -12
timeout 300 python tools/side_bench.py allpairs > gpurun_out/r_side.txt 2>&1; cat gpurun_out/r_side.txt
timeout 300 ncu --set full --clock-control none -k regex:l2max_allpairs -s 2 -c 1 -o gpurun_out/r_allpairs python tools/side_bench.py allpairs > gpurun_out/r_ncu.log 2>&1; tail -2 gpurun_out/r_ncu.log
