#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_l2max_gpu.py -m gpu -q --timeout 120 > gpurun_out/s_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/s_pytest.txt
grep -E "FAIL|passed|failed|exit|Error" gpurun_out/s_pytest.txt | cut -c1-250 | tail -6
timeout 300 python tools/side_bench.py allpairs > gpurun_out/s_side.txt 2>&1; cat gpurun_out/s_side.txt
